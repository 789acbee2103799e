"""Chain split into slab-range pieces on two streams: step time vs number of pieces (dev aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_perf import run

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
plist = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 16, 32]
for pieces in plist:
    os.environ["NSB_CHAIN_PIECES"] = str(pieces)
    print("pieces", pieces, {k: v for k, v in os.environ.items() if k.startswith("NSB_")}, flush=True)
    run(n, 4 if n > 4096 else 20)
