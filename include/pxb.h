/*
 * pxb.h -- C ABI of libpxb, the sm_100a photon-generation / photon-projection hot path.
 *
 * This is the drop-in boundary for pyXSIM's native + NumPy hot path.  Every entry point
 * names the reference interface it replaces (paths relative to the pyXSIM tree):
 *
 *   pxb_model_create ............ SpectralInterpolator1D/2D construction,
 *                                 pyxsim/spectral_models.py:22-62 (tables come from
 *                                 TableCIEModel.prepare_spectrum :253-263 /
 *                                 PionSpectralModel.prepare_spectrum :413-427)
 *   pxb_interp_spec ............. interp1d_spec / interp2d_spec,
 *                                 pyxsim/lib/interpolate.pyx:10-58 / :64-143 (bit-exact)
 *   pxb_thermal_spectra ......... get_spectrum + abundance combine + clip,
 *                                 pyxsim/source_models/thermal_sources/base.py:453-460
 *   pxb_thermal_counts .......... cut mask, cell_nrm, spec_sum, Poisson counts,
 *                                 base.py:335-354,423-425,463-468
 *   pxb_scan_counts ............. the implicit running cursor `ei`/`end_e`, base.py:474-491,
 *                                 and active-cell selection, base.py:517-520
 *   pxb_compact_cells ........... gather/unwrap/recentre of active cells,
 *                                 pyxsim/photon_list.py:423-451
 *   pxb_thermal_sample .......... per-cell CDF + inverse-CDF sampling, base.py:471-491,522-523
 *   pxb_powerlaw_counts/sample .. pyxsim/source_models/power_law_sources.py:186-250
 *   pxb_line_counts/sample ...... pyxsim/source_models/line_sources.py:193-234
 *   pxb_radius2 ................. SourceModel.compute_radius, source_models/sources.py:77-87
 *   pxb_doppler_shift ........... doppler_shift, pyxsim/lib/sky_functions.pyx:20-34
 *   pxb_pixel_to_cel ............ pixel_to_cel, sky_functions.pyx:153-179
 *   pxb_absorb_transmission ..... AbsorptionModel.get_absorb, spectral_models.py:556-561
 *                                 (+ soxs get_wabs_absorb used by WabsModel :633-635)
 *   pxb_thermal_spectrum ........ process_data("spectrum") + shift_spectrum,
 *                                 base.py:453-460,494-495; pyxsim/lib/spectra.pyx:67-93
 *   pxb_thermal_band ............ process_data("*intensity") + make_band, base.py:497-500;
 *                                 spectra.pyx:96-126
 *   pxb_powerlaw_spectrum, pxb_powerlaw_norm ... power_law_spectrum, spectra.pyx:11-32;
 *                                 power_law_sources.py:198-206
 *   pxb_line_spectrum ........... line_spectrum, spectra.pyx:35-64
 *   pxb_shift_factor ............ SourceModel.compute_shift, source_models/sources.py:89-95
 *   pxb_thermal_flux ............ make_fluxf + the "luminosity"/"photon_rate" branch of
 *                                 process_data, spectral_models.py:101-134,464-537; base.py:501-511
 *   pxb_absorb_decide ........... AbsorptionModel.absorb_photons, spectral_models.py:563-583
 *                                 (diagnostic: explicit uniforms)
 *   pxb_cel_to_pixel, pxb_event_image, pxb_event_spectrum ... EventList binning,
 *                                 pyxsim/event_list.py:106-134,254-318,344-349
 *   pxb_column_lookup ........... nH cube interpolation at the rotated cell centres,
 *                                 photon_list.py:672-680 (scipy interpn, linear, fill 0)
 *   pxb_make_events ............. both stages fused: base.py:462-492 -> photon_list.py:686-799
 *   pxb_project ................. the projection tile loop, photon_list.py:686-799:
 *                                 doppler_shift + absorb_photons (spectral_models.py:563-583)
 *                                 + scatter_events (sky_functions.pyx:40-147) or
 *                                 scatter_events_allsky (:185-250) + sigma_pos + /D_A +
 *                                 flat-sky | pixel_to_cel + eobs[det] + flux/emin/emax
 *
 * Conventions
 *   - All array pointers are DEVICE pointers unless a field says "host".
 *   - Sizes are explicit int64 element counts.  `stream` is a cudaStream_t passed as void*.
 *   - Every function returns 0 on success and a negative pxb_status on failure;
 *     pxb_last_error() returns a thread-local message.  No exceptions cross the ABI.
 *   - The caller owns every buffer (two-phase count -> scan -> fill for variable-length
 *     outputs); the only library-owned memory is the table model behind pxb_model.
 *   - Deterministic: outputs are a pure function of (inputs, seed, cell_id0).
 */
#ifndef PXB_H
#define PXB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXB_MAX_VAR 64

enum pxb_status {
  PXB_OK = 0,
  PXB_ERR_INVALID = -1,   /* bad argument */
  PXB_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  PXB_ERR_UNSUPPORTED = -3,
  PXB_ERR_NO_DEVICE = -4
};

const char* pxb_last_error(void);
int pxb_version(void);
/* sm count / compute capability of the current device */
int pxb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* sizeof() of pxb_table_desc, pxb_model_desc, pxb_thermal_cells, pxb_geometry,
 * pxb_powerlaw_cells, pxb_line_cells, pxb_photon_list, pxb_project_params (in that order);
 * returns the number of entries written.  Lets a binding verify its struct layout. */
int pxb_struct_sizes(int64_t* out, int n);
/* Number of kernels libpxb has launched in this process since the last reset (every launch in
 * the library goes through one checked macro that counts it). */
int64_t pxb_launch_count(void);
void pxb_launch_count_reset(void);

/* ------------------------------------------------------------------------------------
 * Thermal spectral tables
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t nx;            /* number of temperature nodes (0 => table absent) */
  int32_t ny;            /* number of density nodes, 0 => 1-D table */
  int32_t nbins;
  int32_t nvar;
  const double* xnodes;  /* host [nx] */
  const double* ynodes;  /* host [ny] or NULL */
  const double* cosmic;  /* host [max(ny,1)*nx, nbins], row = nx*y_i + x_i */
  const double* metal;   /* host, same shape */
  const double* var;     /* host [nvar, rows, nbins] or NULL */
} pxb_table_desc;

typedef struct {
  pxb_table_desc primary;
  pxb_table_desc fallback;    /* 1-D table used outside the primary's (kT,nH) bounds */
  int32_t log_t;              /* primary coordinate is log10(kT*K_per_keV) (y: log10 nH) */
  int32_t fallback_log_t;
  int32_t density_dependent;  /* 2-D lookup, result divided by nH */
  int32_t binscale_log;       /* bin_edges are log10(ebins); sampled energies are 10**E */
  int32_t method;             /* 0 invert_cdf, 1 accept_reject (bin centres) */
  int32_t reserved;
  double K_per_keV;
  double kT_lo, kT_hi, nH_lo, nH_hi; /* bounds of the primary 2-D table (use_igm mask) */
  const double* bin_edges;    /* host [nbins+1] */
  const double* emid;         /* host [nbins] */
  const double* ebins;        /* host [nbins+1] linear bin edges in keV (== bin_edges unless
                                 binscale_log); used by the spectrum / band entry points */
} pxb_model_desc;

typedef struct pxb_model pxb_model;

int pxb_model_create(const pxb_model_desc* desc, pxb_model** out);
int pxb_model_destroy(pxb_model* model);
/* 1 when every table entry is >= 0 (the O(1) row-sum / mixture fast paths are legal) */
int pxb_model_nonnegative(const pxb_model* model);
/* 1 when pxb_make_events is expected to beat pxb_thermal_sample + pxb_project for this model (the
 * common sampler applies: linear uniform bins, invert_cdf, no variable elements, 1-D table); 0:
 * the two calls are faster, although pxb_make_events still gives the same events */
int pxb_model_fusion_pays(const pxb_model* model);

/* Per-chunk inputs of the thermal source (all per-cell arrays are device, length ncell). */
typedef struct {
  int64_t ncell;
  int64_t cell_id0;            /* global id of cell 0: RNG counters use the global id of cell i,
                                  cell_id0 + i, or -- when id_block > 0 (a block-cyclic shard of a
                                  larger dataset gathered into one chunk) --
                                  cell_id0 + (i / id_block) * id_stride + i % id_block */
  const double* kT;            /* keV */
  const double* em;            /* emission measure, cm^-3 */
  const double* nH;            /* H nuclei density cm^-3, or NULL */
  const double* zmet;          /* metallicity field or NULL -> zmet_const */
  const double* xh;            /* hydrogen mass fraction field or NULL -> xh_const */
  const double* density;       /* NULL -> no max_density cut */
  const double* entropy;       /* NULL -> no min_entropy cut */
  const double* r2;            /* cm^2, internal observer (cell_nrm /= r2) or NULL */
  const double* elem[PXB_MAX_VAR];   /* per-element abundance field or NULL -> elem_const[k] */
  double elem_const[PXB_MAX_VAR];
  double elem_scale[PXB_MAX_VAR];    /* mconvert factor (1 when already Zsun) */
  int32_t elem_by_xh[PXB_MAX_VAR];   /* divide the factor by X_H */
  double zmet_const;
  double zmet_scale;           /* Zconvert (1 when already Zsun) */
  int32_t zmet_by_xh;
  int32_t nvar;
  double xh_const;
  double spectral_norm;
  double kT_min, kT_max;
  double max_density, min_entropy;
  uint64_t seed;
  int64_t id_block, id_stride; /* see cell_id0; 0 => contiguous ids */
  int32_t generic_kernels;     /* diagnostic switches, same results either way.  bit 0: never take
                                  the kernels compiled for the common option set (parity tests);
                                  bit 1: pxb_make_events with one big CTA per SM and the table rows
                                  staged in shared memory instead of small CTAs (four per SM) that
                                  gather the rows through L1/L2 */
  int32_t split_factor;        /* variable-element tables: a cell's photons are split over its
                                  (row, table) components once per cell when it holds at least
                                  split_factor * components photons (below that every photon
                                  picks its component from cumulative probabilities stored per
                                  cell); 0 => default (40), < 0 => never, and the per-photon
                                  choice re-derives the weights instead of storing them */
} pxb_thermal_cells;

/* lambda_out (nullable) receives spec_sum*cell_nrm, counts_out the Poisson draws.
 * Cells failing the cut get lambda = 0, count = 0. */
int pxb_thermal_counts(const pxb_model* model, const pxb_thermal_cells* cells,
                       double* lambda_out, int64_t* counts_out, void* stream);

/* Test/diagnostic: clipped total spectrum of every cell, spec_out[ncell, nbins]. */
int pxb_thermal_spectra(const pxb_model* model, const pxb_thermal_cells* cells,
                        double* spec_out, void* stream);

/* Bit-exact replacement of interp1d_spec / interp2d_spec on the model's PRIMARY table.
 * x_vals/y_vals are table coordinates (already log10 when the model is log_t), x_is/y_is
 * the clamped node indices; outputs c,m [n, nbins], v [nvar, n, nbins] (nullable). */
int pxb_interp_spec(const pxb_model* model, int64_t n, const double* x_vals,
                    const int32_t* x_is, const double* y_vals, const int32_t* y_is,
                    double* c_out, double* m_out, double* v_out, void* stream);

/* Inverse-CDF energy sampling for the active cells.  idx[k] is the chunk-local cell index,
 * nph[k] its photon count, poff[k] its first slot in `energies`, qoff[k] its first photon PAIR
 * (both have n_active+1 entries; pxb_scan_counts / pxb_compact_cells produce them).
 * Two device paths with the same output distribution: cells whose tables, abundances and
 * interpolation weights are all non-negative are sampled as a mixture of per-row alias tables
 * (thread per photon pair); every other cell gets its own blended + clipped CDF in shared memory
 * (CTA per cell).  The order of the photons inside a cell is not a contract (the reference
 * sorts its uniforms, base.py:481). */
int pxb_thermal_sample_workspace_bytes(const pxb_model* model, int64_t n_active,
                                       int64_t n_pairs, int64_t* bytes);
int pxb_thermal_sample(const pxb_model* model, const pxb_thermal_cells* cells,
                       int64_t n_active, const int64_t* idx, const int64_t* nph,
                       const int64_t* poff, const int64_t* qoff, int64_t n_photons,
                       int64_t n_pairs, double* energies, void* workspace,
                       int64_t workspace_bytes, void* stream);
/* Diagnostic, host only (no GPU needed): the Walker alias table libpxb builds for one table row
 * (bin j is produced directly with probability thr[j] / 2^32 / nbins and as the alternative of
 * every bin k with other[k] == j), so a CPU test can check the implied bin probabilities. */
int pxb_alias_table_host(const double* row, int nbins, uint32_t* thr_out, uint32_t* other_out);

/* ------------------------------------------------------------------------------------
 * Spectrum / band modes ("next" row f1 of the scope table)
 * ---------------------------------------------------------------------------------- */
/* mode="spectrum" of ThermalSourceModel.process_data (base.py:453-460,494-495) fused with
 * shift_spectrum (pyxsim/lib/spectra.pyx:67-93): spec_out[j] += sum over cells passing the
 * cut of shift^3 * cell_nrm * (C(ebnew[j+1]/shift) - C(ebnew[j]/shift)), C = np.interp of the
 * cell's cumulative clipped spectrum on the model's linear bin edges.  shift is per cell
 * (device) or NULL (= 1).  ebnew: device [nnew+1]; spec_out: device [nnew], ACCUMULATED.
 * Without a shift and with a workspace (pxb_thermal_spectrum_workspace_bytes; may be NULL) the
 * sum is assembled by linearity from per-row weights -- O(1) per cell -- for every cell whose
 * clip is a no-op; the others take the per-cell path. */
int pxb_thermal_spectrum_workspace_bytes(const pxb_model* model, int64_t ncell, int64_t* bytes);
int pxb_thermal_spectrum(const pxb_model* model, const pxb_thermal_cells* cells,
                         const double* shift, int64_t nnew, const double* ebnew,
                         double* spec_out, void* workspace, int64_t workspace_bytes,
                         void* stream);
/* make_band (spectra.pyx:96-126) on the cell's clipped spectrum, times cell_nrm
 * (base.py:497-500): out[i] = cell_nrm * shift^(3 + use_energy) * sum_j ener_j * spec_ij over
 * the bins with ebins[j] >= emin/shift and ebins[j+1] <= emax/shift; 0 for cut cells.  With a
 * workspace (pxb_thermal_band_workspace_bytes; may be NULL) cells whose clip is a no-op are done
 * in O(1) from per-row prefix sums, the others per cell. */
int pxb_thermal_band_workspace_bytes(int64_t ncell, int64_t* bytes);
int pxb_thermal_band(const pxb_model* model, const pxb_thermal_cells* cells, int use_energy,
                     double emin, double emax, const double* shift, double* out, void* workspace,
                     int64_t workspace_bytes, void* stream);
/* Band fluxes of a table, per table row (what make_fluxf tabulates, spectral_models.py:101-134,
 * 464-478): device arrays [nrows] (vflux: [nvar][nrows] or NULL); fb_* describe the 1-D CIE
 * fallback table of a density-dependent model (NULL otherwise). */
typedef struct {
  const double* cflux;
  const double* mflux;
  const double* vflux;
  const double* fb_cflux;
  const double* fb_mflux;
  const double* fb_vflux;
} pxb_band_flux;
/* modes "luminosity" / "photon_rate" of ThermalSourceModel.process_data (base.py:501-511) with
 * the flux function of make_fluxf: out[i] = cell_nrm_i * (cf(T_i) + Z_i mf(T_i) + sum_k e_ik
 * vf_k(T_i)), 0 for cells failing the cut, NOT clipped.  cf/mf/vf: scipy interp1d (linear) of
 * the row fluxes in the table's temperature coordinate (spectral_models.py:109-131); for a
 * density-dependent model the bilinear _get_flux_2d divided by nH inside the (T, nH) table and
 * the CIE fallback outside (spectral_models.py:482-537).  n_outside_out (device int64,
 * ACCUMULATED) counts selected cells whose temperature lies outside the 1-D table -- where
 * scipy's interp1d raises ValueError; their out[i] is NaN. */
int pxb_thermal_flux(const pxb_model* model, const pxb_thermal_cells* cells,
                     const pxb_band_flux* band, double* out, int64_t* n_outside_out, void* stream);
/* power_law_spectrum (spectra.pyx:11-32): spec_out[j] += sum_i shift_i^2 K_i (emid_j/shift_i)^-alpha_i */
int pxb_powerlaw_spectrum(int64_t ncell, const double* alpha, const double* K, const double* shift,
                          int64_t nbins, const double* emid, double* spec_out, void* stream);
/* line_spectrum (spectra.pyx:35-64): spec_out[j] += sum_i shift_i^2 N_i g((ee_j/shift_i - e0)/sigma_i)/sigma_i,
 * g = np.interp on the tabulated unit Gaussian (gx ascending, device [ng]). */
int pxb_line_spectrum(int64_t ncell, double e0, const double* sigma, const double* N,
                      const double* shift, int64_t nbins, const double* ee, int64_t ng,
                      const double* gx, const double* gpdf, double* spec_out, void* stream);

/* Doppler factor per cell for the spectrum modes, SourceModel.compute_shift (sources.py:89-95):
 * sqrt(1 - beta^2) / (1 - beta . n_hat); velocities in km/s, normal = host double[3]. */
int pxb_shift_factor(int64_t n, const double* vx, const double* vy, const double* vz,
                     const double* normal, double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Counts -> offsets, active-cell compaction
 * ---------------------------------------------------------------------------------- */
int pxb_scan_workspace_bytes(int64_t n, int64_t* bytes);
/* offsets[n+1]: exclusive prefix of counts; arank[n+1]: exclusive prefix of (counts>0);
 * pairoff[n+1]: exclusive prefix of ceil(counts/2) (photon pairs: the unit of work of the photon
 * kernels); totals[3] (device) = {sum(counts), number of active cells, number of pairs}. */
int pxb_scan_counts(const int64_t* counts, int64_t n, int64_t* offsets, int64_t* arank,
                    int64_t* pairoff, int64_t* totals, void* workspace, int64_t workspace_bytes,
                    void* stream);

typedef struct {
  double c[3];        /* centre, kpc */
  double le[3], re[3], dw[3]; /* object bounds and domain width, kpc */
  double bulk_v[3];   /* km/s */
  int32_t periodic[3];
  int32_t reserved;
} pxb_geometry;

/* Gather the active cells: idx/nph/poff plus recentred, periodically unwrapped positions
 * (photon_list.py:423-451) and bulk-subtracted velocities.  dx may be NULL (point sources). */
int pxb_compact_cells(int64_t n, const int64_t* counts, const int64_t* offsets,
                      const int64_t* arank, const int64_t* pairoff, const double* x, const double* y,
                      const double* z, const double* vx, const double* vy, const double* vz,
                      const double* dx, const pxb_geometry* geom, int64_t* idx_out,
                      int64_t* nph_out, int64_t* poff_out, int64_t* qoff_out, double* xo, double* yo,
                      double* zo, double* vxo, double* vyo, double* vzo, double* dxo, void* stream);

/* r2[i] = |pos - c|^2 * cm2_per_kpc2 with periodic unwrap (sources.py:77-87). */
int pxb_radius2(int64_t n, const double* x, const double* y, const double* z,
                const pxb_geometry* geom, double* r2_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Power-law and line sources
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t ncell;
  int64_t cell_id0;
  const double* lum;      /* luminosity in keV/s */
  const double* alpha;    /* photon index field or NULL -> alpha_const */
  const double* r2;       /* cm^2 or NULL */
  double alpha_const;
  double e0, emin, emax;  /* keV */
  double spectral_norm;
  double scale_factor;    /* 1/(1+z) */
  uint64_t seed;
  int64_t id_block, id_stride;  /* as in pxb_thermal_cells */
} pxb_powerlaw_cells;

int pxb_powerlaw_counts(const pxb_powerlaw_cells* cells, double* lambda_out,
                        int64_t* counts_out, void* stream);
int pxb_powerlaw_sample(const pxb_powerlaw_cells* cells, int64_t n_active, const int64_t* idx,
                        const int64_t* nph, const int64_t* poff, int64_t n_photons,
                        double* energies, void* stream);

typedef struct {
  int64_t ncell;
  int64_t cell_id0;
  const double* rate;     /* photons/s (cgs value of the emission field) */
  const double* sigma;    /* line width in keV per cell, or NULL -> sigma_const */
  const double* r2;
  double sigma_const;
  double e0;              /* keV */
  double spectral_norm;
  double scale_factor;
  uint64_t seed;
  int64_t id_block, id_stride;  /* as in pxb_thermal_cells */
} pxb_line_cells;

/* K = L / K_fac and the per-cell alpha (power_law_sources.py:196-206), inputs of the spectrum mode */
int pxb_powerlaw_norm(const pxb_powerlaw_cells* cells, double* K_out, double* alpha_out, void* stream);

int pxb_line_counts(const pxb_line_cells* cells, double* lambda_out, int64_t* counts_out,
                    void* stream);
int pxb_line_sample(const pxb_line_cells* cells, int64_t n_active, const int64_t* idx,
                    const int64_t* nph, const int64_t* poff, int64_t n_photons,
                    double* energies, void* stream);

/* Diagnostics: `count` independent draws of the device's Poisson / binomial samplers (the
 * counter-based replacements of prng.poisson, base.py:466, and of the per-photon component
 * choice); element i uses the Philox stream of "cell" i. */
int pxb_rng_poisson(double lam, uint64_t seed, int64_t count, int64_t* out, void* stream);
int pxb_rng_binomial(int64_t n, double p, uint64_t seed, int64_t count, int64_t* out,
                     void* stream);

/* ------------------------------------------------------------------------------------
 * Projection
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t n_cells;
  int64_t n_photons;
  int64_t n_pairs;       /* sum over cells of ceil(nph / 2) == qoff[n_cells] */
  const int64_t* nph;    /* [n_cells] (may be NULL: poff carries the counts) */
  const int64_t* poff;   /* [n_cells+1] exclusive offsets into energy */
  const int64_t* qoff;   /* [n_cells+1] exclusive prefix of ceil(nph / 2): the photon kernels walk
                            the photons of a cell in pairs (pxb_scan_counts produces it) */
  const double* x;       /* kpc, centred */
  const double* y;
  const double* z;
  const double* vx;      /* km/s, bulk-subtracted */
  const double* vy;
  const double* vz;
  const double* dx;      /* kpc (cell width or smoothing length) */
  const double* energy;  /* [n_photons] keV (NULL for pxb_make_events) */
  const int64_t* cell_id; /* [n_cells] global id of every active cell (keys the projection's RNG
                             counters, so the events of a cell do not depend on how the dataset was
                             chunked or sharded over ranks), or NULL -> params.cell_id0 + c */
} pxb_photon_list;

enum { PXB_ABS_NONE = 0, PXB_ABS_TABLE = 1, PXB_ABS_WABS = 2 };
enum { PXB_SKY_TAN = 0, PXB_SKY_FLAT = 1, PXB_SKY_PHYS = 2 };

typedef struct {
  int32_t observer;      /* 0 external, 1 internal (all-sky) */
  int32_t data_type;     /* 0 cells, 1 particles */
  int32_t kernel;        /* 0 top_hat, 1 gaussian */
  int32_t no_shifting;
  int32_t vn_axis;       /* 0/1/2: that velocity component; -1: v . z_hat */
  int32_t sky_mode;      /* PXB_SKY_* */
  int32_t save_los;
  int32_t absorb_kind;   /* PXB_ABS_* */
  int32_t has_sigma_pos;
  int32_t events_f32;    /* 1: xsky / ysky / eobs / los are FLOAT arrays (arithmetic stays fp64): eobs
                            rounded to single precision, xsky / ysky as offsets from sky_center
                            for the angular sky modes of an external observer (absolute values
                            otherwise) -- half the bytes of an event list */
  double sigma_pos;
  double x_hat[3], y_hat[3], z_hat[3];
  double D_A_kpc;
  double sky_center[2];  /* degrees */
  double nH;             /* foreground column, 1e22 cm^-2 */
  const double* abs_energy;  /* device [abs_n], ascending, keV */
  const double* abs_sigma;   /* device [abs_n], in 1e-22 cm^2 so that tau = sigma * nH */
  int64_t abs_n;
  int32_t abs_grid;          /* 0 arbitrary ascending grid (binary search), 1 uniform in E,
                                2 uniform in log10(E): index = (x - abs_x0) * abs_inv_dx */
  int32_t generic_kernels;   /* diagnostic: never take the kernels compiled for the common
                                option set (same events; used by the parity tests) */
  double abs_x0, abs_inv_dx;
  const double* nH_cell;     /* device per-cell internal column (1e22 cm^-2) or NULL */
  double oneplusz;
  uint64_t seed;
  int64_t cell_id0;
} pxb_project_params;

int pxb_project_workspace_bytes(int64_t n_cells, int64_t n_pairs, int64_t* bytes);
/* Outputs xsky/ysky/eobs (and los when save_los) must hold n_photons entries (doubles, or floats
 * with params->events_f32); the first n_events are valid and keep the photon-list order (stable
 * compaction).
 * n_events_out: device int64[1].  stats_out: device double[3] = {sum(eobs), min, max}. */
int pxb_project(const pxb_photon_list* photons, const pxb_project_params* params,
                double* xsky, double* ysky, double* eobs, double* los,
                int64_t* n_events_out, double* stats_out, void* workspace,
                int64_t workspace_bytes, void* stream);

/* Generation and projection of a thermal source in ONE pass, without a photon list in between
 * (the two reference stages base.py:462-492 -> photon_list.py:686-799 back to back): `photons`
 * describes the active cells as pxb_compact_cells wrote them (energy = NULL), idx their
 * chunk-local indices, cells/model the thermal source as for pxb_thermal_sample.  The events are
 * bit for bit those of pxb_thermal_sample followed by pxb_project with the same seeds.
 * n_declined_out (device int64[1]) receives the number of active cells the mixture sampler cannot
 * take (extrapolating beyond the table, negative abundance): they produce NO events here, so a
 * caller must fall back to the two-stage path when it is not zero.  Needs non-negative tables. */
int pxb_make_events_workspace_bytes(const pxb_model* model, int64_t n_cells, int64_t n_pairs,
                                    int64_t* bytes);
int pxb_make_events(const pxb_model* model, const pxb_thermal_cells* cells, const int64_t* idx,
                    const pxb_photon_list* photons, const pxb_project_params* params,
                    double* xsky, double* ysky, double* eobs, double* los, int64_t* n_events_out,
                    double* stats_out, int64_t* n_declined_out, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* Diagnostic: the variant of the photon kernel the calling thread's last pxb_thermal_sample /
 * pxb_project / pxb_make_events launched: mode 0 sample, 1 project, 2 fused; spec 0 = option-generic
 * kernel, 1 / 2 = compiled for the common option set (on- / off-axis); staged = table rows kept in
 * shared memory.  Returns PXB_ERR_INVALID before the first launch. */
int pxb_last_pipeline_launch(int* mode, int* spec, int* staged);

/* In-place Doppler shift with per-cell factors (sky_functions.pyx:20-34). */
int pxb_doppler_shift(int64_t n_cells, const double* beta_n, const double* beta2,
                      const int64_t* poff, int64_t n_photons, double* eobs, void* stream);
/* In-place inverse gnomonic projection (sky_functions.pyx:153-179); radians in, degrees out. */
int pxb_pixel_to_cel(int64_t n, double* xsky, double* ysky, double ra0_deg, double dec0_deg,
                     void* stream);
/* transmission exp(-sigma(E)*nH) for n energies with the absorption described by params
 * (absorb_kind, nH, abs_*). */
int pxb_absorb_transmission(const pxb_project_params* params, int64_t n, const double* energy,
                            double* trans_out, void* stream);
/* Internal-absorption columns (photon_list.py:672-680): rotate every cell centre into the
 * projection frame, pos = (r.x_hat, r.y_hat, r.z_hat) with unit_vectors[9] row-major
 * (east, north, normal), and interpolate the column-density cube nH_grid[nw][nw][nd] (C order)
 * trilinearly at pos on the node arrays wmid[nw] (both sky axes) and dmid[nd] (depth):
 * scipy.interpolate.interpn(..., method="linear", bounds_error=False, fill_value=0.0).  The
 * result is what pxb_project_params.nH_cell expects. */
int pxb_column_lookup(int64_t n_cells, const double* x, const double* y, const double* z,
                      const double* unit_vectors, int64_t nw, int64_t nd, const double* wmid,
                      const double* dmid, const double* nH_grid, double* nH_cell_out,
                      void* stream);
/* Diagnostic: the projection kernels' own absorption decision u < exp(-sigma(E)*nH) for n
 * caller-supplied (energy, u) pairs (AbsorptionModel.absorb_photons, spectral_models.py:563-583,
 * with the uniforms made explicit); survive_out[i] = 1 if photon i is detected.  Lets a test put
 * u on either side of the threshold and E on the wabs piece edges. */
int pxb_absorb_decide(const pxb_project_params* params, int64_t n, const double* energy,
                      const double* u, uint8_t* survive_out, void* stream);
/* Diagnostic: the decision as the photon kernels take it, from random WORDS.  The absorption
 * uniform of photon i is U = (w[i] + w2 / 2^32) / 2^32, where w2 -- drawn only when w[i] alone
 * does not decide -- is word 0 of Philox4x32-10(key = params->seed, counter = (draw0 + i,
 * cell_key, stream 10)); survive_out[i] = (U < exp(-sigma(E_i) nH)). */
int pxb_absorb_decide_words(const pxb_project_params* params, int64_t n, const double* energy,
                            const uint32_t* w, uint64_t cell_key, uint64_t draw0,
                            uint8_t* survive_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Event-list consumers (pyxsim/event_list.py)
 * ---------------------------------------------------------------------------------- */
/* wcs_world2pix (origin 1) for the TAN header EventList builds (event_list.py:106-112,281-287):
 * CRVAL = (ra0, dec0), CRPIX = crpix on both axes, CDELT = (cdelt1, cdelt2) in degrees.  Points
 * on the far hemisphere give NaN. */
int pxb_cel_to_pixel(int64_t n, const double* ra_deg, const double* dec_deg, double ra0_deg,
                     double dec0_deg, double crpix, double cdelt1, double cdelt2,
                     double* xpix_out, double* ypix_out, void* stream);
/* write_fits_image's binning (event_list.py:254-318): counts of the events with
 * emin <= eobs <= emax per pixel of an nx x nx TAN image of pixel size dtheta_deg centred on
 * (ra0, dec0); np.histogram2d edge semantics; image[iy*nx + ix] (the H.T the reference writes),
 * uint64, ACCUMULATED. */
int pxb_event_image(int64_t n, const double* xsky, const double* ysky, const double* eobs,
                    double emin, double emax, double ra0_deg, double dec0_deg, int nx,
                    double dtheta_deg, uint64_t* image, void* stream);
/* write_spectrum's binning (event_list.py:344-349): np.histogram(eobs, bins=edges), edges device
 * [nchan+1] ascending; counts uint64 [nchan], ACCUMULATED. */
int pxb_event_spectrum(int64_t n, const double* eobs, int nchan, const double* edges,
                       uint64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PXB_H */
