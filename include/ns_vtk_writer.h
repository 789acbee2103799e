/*
 * include/ns_vtk_writer.h -- VTK export of the spectral solver's fields, B200 backend.
 *
 * The spectral half of /root/reference/src/output/vtk_writer.h (:22-28 and the PVD index :30-32; the
 * finite-difference writers belong to the FD solver, which is out of scope): same signatures, byte-identical
 * files. The reference's functions write whatever the host arrays state->x, y, u, v, omega hold; so do these.
 * ns_spectral_write_vtk_strided is additive: it first refreshes u, v, omega from the device (needed under
 * NS_SYNC_LAZY) and writes every stride-th point (a full ASCII dump of 16384^2 points would be ~15 GB).
 */
#ifndef NS_VTK_WRITER_H
#define NS_VTK_WRITER_H

#include "ns_spectral_solver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* vtk_writer.c:212-227 : legacy ASCII STRUCTURED_GRID, vectors "velocity", scalars "vorticity" */
int ns_vtk_write_spectral_legacy(const char *filename, const NSSpectralData *state, double time);
/* vtk_writer.c:229-244 : XML RectilinearGrid (.vtr), ASCII */
int ns_vtk_write_spectral_xml(const char *filename, const NSSpectralData *state, double time);
/* vtk_writer.c:276-320 : create / append to a ParaView .pvd collection */
int ns_vtk_update_pvd(const char *pvd_filename, const char *dataset_file, double time);

/* additive: sync u, v, omega from the device, then the legacy format thinned by `stride` (>= 1) in x and y */
int ns_spectral_write_vtk_strided(NSSpectralData *data, const char *filename, int stride);

#ifdef __cplusplus
}
#endif
#endif
