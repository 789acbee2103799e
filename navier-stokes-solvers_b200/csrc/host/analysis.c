/*
 * analysis.c -- turbulence statistics and VTK export for the spectral solver:
 * include/ns_statistics.h, include/ns_vtk_writer.h. Host C over the kernel C-ABI; the grid-sized sums of
 * the statistics run on the device, the files are written from the host arrays.
 */
#include "ns_statistics.h"
#include "ns_vtk_writer.h"
#include "nsb.h"

/* spectral_api.c */
void ns_spectral_set_error_text(const char *text);

/* statistics.c:77-102 */
static double integral_scale_from_corr(const double *corr, int len, double spacing) {
    if (!corr || len < 2 || spacing <= 0.0) return 0.0;
    double integral = 0.0;
    for (int s = 0; s < len - 1; ++s) {
        const double c0 = corr[s], c1 = corr[s + 1];
        if (c0 <= 0.0) break;
        if (c1 <= 0.0) {
            const double frac = c0 / (c0 - c1);
            integral += 0.5 * c0 * frac * spacing;
            break;
        }
        integral += 0.5 * (c0 + c1) * spacing;
    }
    return integral;
}

/* statistics.c:241-262 */
static void moments_from_sums(NSFieldMoments *m, double sum2, double sum3, double sum4, double inv_total) {
    const double var = sum2 * inv_total;
    m->rms = (var > 0.0) ? sqrt(var) : 0.0;
    if (m->rms > 0.0) {
        const double inv_rms3 = 1.0 / (m->rms * m->rms * m->rms);
        const double inv_rms4 = inv_rms3 / m->rms;
        m->skewness = (sum3 * inv_total) * inv_rms3;
        m->kurtosis = (sum4 * inv_total) * inv_rms4;
    } else {
        m->skewness = 0.0;
        m->kurtosis = 0.0;
    }
}

/* raw device sums -> the reference's outputs (statistics.c:203-204, 238-336) */
static int finish_statistics(const double raw[10], const double *cx_raw, const double *cy_raw, int nx, int ny, double dx,
                             double dy, double nu, int max_sep_x, int max_sep_y, int sep_x, int sep_y,
                             NSTurbulenceStatistics *stats, double *corr_x, double *corr_y, double *tmp_x, double *tmp_y) {
    const double inv_total = 1.0 / (double)((size_t)nx * (size_t)ny);
    stats->u.mean = raw[0] * inv_total;
    stats->v.mean = raw[1] * inv_total;
    moments_from_sums(&stats->u, raw[2], raw[3], raw[4], inv_total);
    moments_from_sums(&stats->v, raw[5], raw[6], raw[7], inv_total);
    const double var_u = raw[2] * inv_total;
    double *bx = corr_x ? corr_x : tmp_x, *by = corr_y ? corr_y : tmp_y;
    /* statistics.c:10-41 / :43-75 : variance <= 0 gives the delta function; norm = 1 / (N var) */
    for (int s = 0; s < sep_x; ++s) bx[s] = (var_u > 0.0) ? cx_raw[s] * (1.0 / ((double)((size_t)nx * ny) * var_u)) : (s == 0 ? 1.0 : 0.0);
    for (int s = 0; s < sep_y; ++s) by[s] = (var_u > 0.0) ? cy_raw[s] * (1.0 / ((double)((size_t)nx * ny) * var_u)) : (s == 0 ? 1.0 : 0.0);
    if (corr_x) for (int s = sep_x; s < max_sep_x; ++s) corr_x[s] = (sep_x == 0 && s == 0) ? 1.0 : 0.0;
    if (corr_y) for (int s = sep_y; s < max_sep_y; ++s) corr_y[s] = (sep_y == 0 && s == 0) ? 1.0 : 0.0;
    stats->integral_length_scale_x = (var_u > 0.0 && sep_x > 1) ? integral_scale_from_corr(bx, sep_x, dx) : 0.0;
    stats->integral_length_scale_y = (var_u > 0.0 && sep_y > 1) ? integral_scale_from_corr(by, sep_y, dy) : 0.0;
    /* statistics.c:104-130 */
    stats->taylor_microscale = 0.0;
    if (nx >= 3 && var_u > 0.0 && dx > 0.0) {
        const double avg = raw[8] * inv_total;
        if (avg > 0.0) stats->taylor_microscale = sqrt(var_u / avg);
    }
    /* statistics.c:132-175 */
    stats->kolmogorov_scale = 0.0;
    if (nx >= 2 && ny >= 2 && nu > 0.0) {
        const double epsilon = 2.0 * nu * (raw[9] * inv_total);
        if (epsilon > 0.0) stats->kolmogorov_scale = pow((nu * nu * nu) / epsilon, 0.25);
    }
    return 0;
}

static int run_statistics(nsb_ctx *ctx, const double *u, const double *v, int nx, int ny, double dx, double dy, double nu,
                          int max_sep_x, int max_sep_y, NSTurbulenceStatistics *stats, double *corr_x, double *corr_y) {
    const int sep_x = (max_sep_x > nx) ? nx : max_sep_x, sep_y = (max_sep_y > ny) ? ny : max_sep_y;
    double raw[10];
    double *scr = (double *)calloc((size_t)(2 * (sep_x + sep_y) + 4), sizeof(double));
    if (!scr) return -1;
    double *cx_raw = scr, *cy_raw = scr + sep_x, *tmp_x = cy_raw + sep_y, *tmp_y = tmp_x + sep_x;
    int rc;
    if (ctx) {
        rc = nsb_velocity_statistics(ctx, dx, dy, sep_x, sep_y, raw, cx_raw, cy_raw);
    } else {
        const char *env = getenv("NS_SPECTRAL_DEVICE");
        rc = nsb_velocity_statistics_host(env ? atoi(env) : 0, u, v, nx, ny, dx, dy, sep_x, sep_y, raw, cx_raw, cy_raw);
    }
    if (rc != 0) {
        ns_spectral_set_error_text(nsb_last_error());
        free(scr);
        return -1;
    }
    rc = finish_statistics(raw, cx_raw, cy_raw, nx, ny, dx, dy, nu, max_sep_x, max_sep_y, sep_x, sep_y, stats, corr_x, corr_y,
                           tmp_x, tmp_y);
    free(scr);
    return rc;
}

int ns_stats_compute_velocity_statistics(const double *NS_RESTRICT u, const double *NS_RESTRICT v, int nx, int ny,
                                         double dx, double dy, double nu, int max_sep_x, int max_sep_y,
                                         NSTurbulenceStatistics *stats, double *corr_x, double *corr_y) {
    if (!u || !v || !stats || nx <= 0 || ny <= 0 || dx <= 0.0 || dy <= 0.0 || max_sep_x < 0 || max_sep_y < 0) return -1;
    return run_statistics(NULL, u, v, nx, ny, dx, dy, nu, max_sep_x, max_sep_y, stats, corr_x, corr_y);
}

int ns_spectral_turbulence_statistics(NSSpectralData *d, int max_sep_x, int max_sep_y, NSTurbulenceStatistics *stats,
                                      double *corr_x, double *corr_y) {
    if (!d || !stats || max_sep_x < 0 || max_sep_y < 0) return -1;
    const double dx = d->Lx / d->nx, dy = d->Ly / d->ny, nu = 1.0 / d->Re;
    if (ns_spectral_get_sync_policy(d) == NS_SYNC_LAZY) {
        /* device resident: u, v from omega_hat (spectral_core.c:37-43), sums on the device, nothing downloaded */
        ns_spectral_compute_derivatives(d);
        nsb_ctx *ctx = (nsb_ctx *)ns_spectral_backend_context(d);
        if (!ctx) return -1;
        return run_statistics(ctx, NULL, NULL, d->nx, d->ny, dx, dy, nu, max_sep_x, max_sep_y, stats, corr_x, corr_y);
    }
    return run_statistics(NULL, d->u, d->v, d->nx, d->ny, dx, dy, nu, max_sep_x, max_sep_y, stats, corr_x, corr_y);
}

/* ------------------------------------------------------------------------ VTK */

/* vtk_writer.c:10-62 with every `stride`-th point (stride 1: the reference's file byte for byte) */
static int write_legacy(const char *filename, const NSSpectralData *s, double time, int stride) {
    FILE *fp = fopen(filename, "w");
    if (!fp) return -1;
    const int nxs = (s->nx + stride - 1) / stride, nys = (s->ny + stride - 1) / stride;
    const long points = (long)nxs * nys;
    fprintf(fp, "# vtk DataFile Version 3.0\n");
    fprintf(fp, "Navier-Stokes Solution t=%.6f\n", time);
    fprintf(fp, "ASCII\n");
    fprintf(fp, "DATASET STRUCTURED_GRID\n");
    fprintf(fp, "DIMENSIONS %d %d 1\n", nxs, nys);
    fprintf(fp, "POINTS %ld double\n", points);
    for (int j = 0; j < s->ny; j += stride)
        for (int i = 0; i < s->nx; i += stride) fprintf(fp, "%.9f %.9f 0.0\n", s->x[i], s->y[j]);
    fprintf(fp, "FIELD FieldData 1\n");
    fprintf(fp, "TIME 1 1 double\n%.9f\n", time);
    fprintf(fp, "POINT_DATA %ld\n", points);
    if (s->u && s->v) {
        fprintf(fp, "VECTORS %s double\n", "velocity");
        for (int j = 0; j < s->ny; j += stride)
            for (int i = 0; i < s->nx; i += stride) {
                const size_t idx = (size_t)j * s->nx + i;
                fprintf(fp, "%.9f %.9f 0.0\n", s->u[idx], s->v[idx]);
            }
    }
    if (s->omega) {
        fprintf(fp, "SCALARS %s double 1\n", "vorticity");
        fprintf(fp, "LOOKUP_TABLE default\n");
        for (int j = 0; j < s->ny; j += stride)
            for (int i = 0; i < s->nx; i += stride) fprintf(fp, "%.9f\n", s->omega[(size_t)j * s->nx + i]);
    }
    return fclose(fp) == 0 ? 0 : -1;
}

int ns_vtk_write_spectral_legacy(const char *filename, const NSSpectralData *state, double time) {
    if (!filename || !state) return -1;
    return write_legacy(filename, state, time, 1);
}

int ns_spectral_write_vtk_strided(NSSpectralData *d, const char *filename, int stride) {
    if (!d || !filename || stride < 1) return -1;
    if (ns_spectral_sync_to_host(d, NS_FIELD_U | NS_FIELD_V | NS_FIELD_OMEGA) != 0) return -1;
    return write_legacy(filename, d, d->t, stride);
}

/* vtk_writer.c:64-141 */
int ns_vtk_write_spectral_xml(const char *filename, const NSSpectralData *s, double time) {
    if (!filename || !s) return -1;
    FILE *fp = fopen(filename, "w");
    if (!fp) return -1;
    const int nx = s->nx, ny = s->ny;
    fprintf(fp, "<?xml version=\"1.0\"?>\n");
    fprintf(fp, "<VTKFile type=\"RectilinearGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
    fprintf(fp, "  <RectilinearGrid WholeExtent=\"0 %d 0 %d 0 0\">\n", nx - 1, ny - 1);
    fprintf(fp, "    <FieldData>\n");
    fprintf(fp, "      <DataArray type=\"Float64\" Name=\"TimeValue\" NumberOfTuples=\"1\" format=\"ascii\">%.9f</DataArray>\n", time);
    fprintf(fp, "    </FieldData>\n");
    fprintf(fp, "    <Piece Extent=\"0 %d 0 %d 0 0\">\n", nx - 1, ny - 1);
    fprintf(fp, "      <PointData");
    if (s->omega) fprintf(fp, " Scalars=\"%s\"", "vorticity");
    if (s->u && s->v) fprintf(fp, " Vectors=\"%s\"", "velocity");
    fprintf(fp, ">\n");
    if (s->omega) {
        fprintf(fp, "        <DataArray type=\"Float64\" Name=\"%s\" format=\"ascii\">\n", "vorticity");
        for (int j = 0; j < ny; ++j) {
            for (int i = 0; i < nx; ++i) fprintf(fp, " %.9f", s->omega[(size_t)j * nx + i]);
            fprintf(fp, "\n");
        }
        fprintf(fp, "        </DataArray>\n");
    }
    if (s->u && s->v) {
        fprintf(fp, "        <DataArray type=\"Float64\" NumberOfComponents=\"3\" Name=\"%s\" format=\"ascii\">\n", "velocity");
        for (int j = 0; j < ny; ++j) {
            for (int i = 0; i < nx; ++i) {
                const size_t idx = (size_t)j * nx + i;
                fprintf(fp, " %.9f %.9f 0.0", s->u[idx], s->v[idx]);
            }
            fprintf(fp, "\n");
        }
        fprintf(fp, "        </DataArray>\n");
    }
    fprintf(fp, "      </PointData>\n");
    fprintf(fp, "      <CellData/>\n");
    fprintf(fp, "      <Coordinates>\n");
    fprintf(fp, "        <DataArray type=\"Float64\" Name=\"X\" format=\"ascii\">\n");
    for (int i = 0; i < nx; ++i) fprintf(fp, " %.9f", s->x[i]);
    fprintf(fp, "\n        </DataArray>\n");
    fprintf(fp, "        <DataArray type=\"Float64\" Name=\"Y\" format=\"ascii\">\n");
    for (int j = 0; j < ny; ++j) fprintf(fp, " %.9f", s->y[j]);
    fprintf(fp, "\n        </DataArray>\n");
    fprintf(fp, "        <DataArray type=\"Float64\" Name=\"Z\" format=\"ascii\"> 0.0 </DataArray>\n");
    fprintf(fp, "      </Coordinates>\n");
    fprintf(fp, "    </Piece>\n");
    fprintf(fp, "  </RectilinearGrid>\n");
    fprintf(fp, "</VTKFile>\n");
    return fclose(fp) == 0 ? 0 : -1;
}

/* vtk_writer.c:276-320 : the index is rewritten with one more <DataSet> line in front of </Collection> */
int ns_vtk_update_pvd(const char *pvd_filename, const char *dataset_file, double time) {
    if (!pvd_filename || !dataset_file) return -1;
    char *old = NULL;
    size_t len = 0;
    FILE *in = fopen(pvd_filename, "r");
    if (in) {
        if (fseek(in, 0, SEEK_END) == 0) {
            const long sz = ftell(in);
            rewind(in);
            if (sz >= 0 && (old = (char *)malloc((size_t)sz + 1)) != NULL) {
                len = fread(old, 1, (size_t)sz, in);
                old[len] = '\0';
            }
        }
        fclose(in);
        if (!old) return -1;
    }
    const char *tail = old ? strstr(old, "</Collection>") : NULL;
    if (old && !tail) {
        free(old);
        return -1;
    }
    FILE *fp = fopen(pvd_filename, "w");
    if (!fp) {
        free(old);
        return -1;
    }
    if (old) {
        fwrite(old, 1, (size_t)(tail - old), fp);
    } else {
        fprintf(fp, "<?xml version=\"1.0\"?>\n"
                    "<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
                    "  <Collection>\n");
    }
    fprintf(fp, "    <DataSet timestep=\"%.9f\" group=\"\" part=\"0\" file=\"%s\"/>\n", time, dataset_file);
    fputs(old ? tail : "  </Collection>\n</VTKFile>\n", fp);
    free(old);
    return fclose(fp) == 0 ? 0 : -1;
}
