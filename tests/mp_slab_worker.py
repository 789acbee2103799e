"""Worker of tests/test_multiprocess_slab.py: one rank of a world_size-2 gloo group on CPU.
Checks the host-side logic of the slab decomposition that needs more than one process: the
partitions of the ranks tile the retained columns / rows exactly, every rank replays the same
seeded forcing stream bit for bit, and 192-byte handle blobs survive the all-gather used by
bench.py. (The kernels themselves run under the single-process emulator tests; CUDA IPC
needs real devices and is exercised by `bench.py --gpus N` on the GPU box.)"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import hostlib  # noqa: E402
import orc  # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nsb = ge.load_package()
    lib = nsb.load_nsb(os.path.join(ROOT, "tests", "emu", "build", "libnsb_emu.so"))
    nx, ny = 1024, 2048
    c = nsb.NsbContext(nx, ny, Re=500.0, lib=lib)
    c.mgpu_setup(rank, world)
    p = c.partition()
    mine = torch.tensor([p["col0"], p["col1"], p["row0"], p["rows"]], dtype=torch.int64)
    allp = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allp, mine)
    cols = sorted((int(t[0]), int(t[1])) for t in allp)
    rows = sorted((int(t[2]), int(t[3])) for t in allp)
    ncol = (2 * (nx // 2 + 1)) // 3 + 1
    assert cols[0][0] == 0 and cols[-1][1] == ncol and all(cols[k][1] == cols[k + 1][0] for k in range(world - 1)), cols
    assert rows[0][0] == 0 and all(rows[k][0] + rows[k][1] == rows[k + 1][0] for k in range(world - 1)), rows
    assert sum(r[1] for r in rows) == ny
    # CUDA IPC is refused loudly under the emulator (no silent fallback)
    try:
        c.mgpu_export()
        raise AssertionError("export should fail under the emulator")
    except nsb.NsbError as e:
        assert "IPC" in str(e)
    # the blob exchange bench.py uses
    blob = torch.tensor(list(bytes([rank]) * 192), dtype=torch.uint8)
    got = [torch.zeros(192, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(got, blob)
    assert all(bytes(got[r].tolist()) == bytes([r]) * 192 for r in range(world))
    # every rank replays the same forcing stream
    o = orc.Oracle("port", 64, 64, Re=100.0)
    o.init_forcing("white", seed=42)
    h = hashlib.sha256()
    for _ in range(3):
        h.update(np.ascontiguousarray(o.gen_forcing(1e-3)).tobytes())
    o.close()
    dig = torch.tensor(list(h.digest()), dtype=torch.uint8)
    digs = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(digs, dig)
    assert all(torch.equal(d, digs[0]) for d in digs)
    c.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
