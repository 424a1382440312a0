/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * clover_oracle.c: a plain-C, single-file restatement of the arithmetic of
 * CloverLeaf_ref's `use_c_kernels` kernel layer (CloverLeaf_ref/kernels/ *_kernel_c.c),
 * exporting the same Fortran-callable symbols (lowercase + trailing underscore, every
 * argument by reference) so the same host driver and the same per-kernel A/B tests can
 * run it, the reference's own objects (oracle/_ref/libclover_ref_c.so) and the CUDA
 * library (cloverleaf_b200/libclover_b200.so) interchangeably.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file
 *   (a) bit-for-bit against oracle/_ref (the reference sources compiled where they lie,
 *       -ffp-contract=off) on random inputs for every entry point and on whole runs, and
 *   (b) against the reference's golden kinetic-energy constants
 *       (CloverLeaf_ref/field_summary.f90:139-143) and the committed traces in tests/golden/.
 *
 * The evaluation ORDER of every floating-point expression follows the reference's C source
 * (left-to-right, no FMA contraction: build with -ffp-contract=off), because the CUDA
 * kernels are built -fmad=false and are expected to match these results bit for bit.
 *
 * Layout (CloverLeaf_ref/kernels/ftocmacros.h:14, build_field.f90:33-94): Fortran
 * column-major arrays with lower bound -1 in both dimensions; element (j,k) of an array
 * with row length R lives at (k+1)*R + (j+1).  R = nx+4 for cell-centred and y-face
 * arrays, nx+5 for vertex, x-face and work arrays.  x_min = y_min = 1 always
 * (start.f90:77-80) but the bounds are honoured as passed.
 */
#include <math.h>
#include <stdio.h>
#include <sys/time.h>

#define DMAX(a, b) ((a) >= (b) ? (a) : (b))
#define DMIN(a, b) ((a) >= (b) ? (b) : (a))

typedef struct {
  int x0, x1, y0, y1; /* x_min, x_max, y_min, y_max */
  int rc, rv;         /* row length of cell/y-face arrays, of vertex/x-face/work arrays */
} grid_t;

static grid_t make_grid(const int *xmin, const int *xmax, const int *ymin, const int *ymax) {
  grid_t g;
  g.x0 = *xmin; g.x1 = *xmax; g.y0 = *ymin; g.y1 = *ymax;
  g.rc = g.x1 + 4;
  g.rv = g.x1 + 5;
  return g;
}
/* index helpers; lower bounds are x_min-2 / y_min-2 */
#define IC(j, k) ((size_t)((k) - (g.y0 - 2)) * g.rc + ((j) - (g.x0 - 2)))
#define IV(j, k) ((size_t)((k) - (g.y0 - 2)) * g.rv + ((j) - (g.x0 - 2)))
#define I1X(j) ((j) - (g.x0 - 2))
#define I1Y(k) ((k) - (g.y0 - 2))

/* ------------------------------------------------------------------------------------ */
/* ideal_gas_kernel_c.c:30-63 */
void ideal_gas_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *density,
                         double *energy, double *pressure, double *soundspeed) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const double gm1 = 1.4 - 1.0;
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      const size_t c = IC(j, k);
      const double rho = density[c];
      const double v = 1.0 / rho;
      const double p = gm1 * rho * energy[c];
      pressure[c] = p;
      const double pe = gm1 * rho;
      const double pv = -rho * p;
      const double ss2 = v * v * (p * pe - pv);
      soundspeed[c] = sqrt(ss2);
    }
}

/* ------------------------------------------------------------------------------------ */
/* viscosity_kernel_c.c:31-110 */
void viscosity_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *celldx, double *celldy,
                         double *density0, double *pressure, double *viscosity, double *xvel0,
                         double *yvel0) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      const double u00 = xvel0[IV(j, k)], u10 = xvel0[IV(j + 1, k)];
      const double u01 = xvel0[IV(j, k + 1)], u11 = xvel0[IV(j + 1, k + 1)];
      const double v00 = yvel0[IV(j, k)], v10 = yvel0[IV(j + 1, k)];
      const double v01 = yvel0[IV(j, k + 1)], v11 = yvel0[IV(j + 1, k + 1)];
      const double dx = celldx[I1X(j)], dy = celldy[I1Y(k)];
      const double ugrad = (u10 + u11) - (u00 + u01);
      const double vgrad = (v01 + v11) - (v00 + v10);
      const double div = dx * ugrad + dy * vgrad;
      const double strain2 =
          0.5 * (u01 + u11 - u00 - u10) / dy + 0.5 * (v10 + v11 - v00 - v01) / dx;
      double pgradx = (pressure[IC(j + 1, k)] - pressure[IC(j - 1, k)]) / (dx + celldx[I1X(j + 1)]);
      double pgrady = (pressure[IC(j, k + 1)] - pressure[IC(j, k - 1)]) / (dy + celldy[I1Y(k + 1)]);
      const double pgradx2 = pgradx * pgradx, pgrady2 = pgrady * pgrady;
      const double limiter =
          ((0.5 * ugrad / dx) * pgradx2 + (0.5 * vgrad / dy) * pgrady2 + strain2 * pgradx * pgrady) /
          DMAX(pgradx2 + pgrady2, 1.0e-16);
      if (limiter > 0.0 || div >= 0.0) {
        viscosity[IC(j, k)] = 0.0;
      } else {
        /* SIGN(MAX(1e-16,|p|),p): magnitude is positive, so only p<0 flips */
        double ax = DMAX(1.0e-16, fabs(pgradx)), ay = DMAX(1.0e-16, fabs(pgrady));
        pgradx = (pgradx < 0.0) ? -ax : ax;
        pgrady = (pgrady < 0.0) ? -ay : ay;
        const double pgrad = sqrt(pgradx * pgradx + pgrady * pgrady);
        const double xgrad = fabs(dx * pgrad / pgradx);
        const double ygrad = fabs(dy * pgrad / pgrady);
        const double grad = DMIN(xgrad, ygrad);
        const double grad2 = grad * grad;
        viscosity[IC(j, k)] = 2.0 * density0[IC(j, k)] * grad2 * limiter * limiter;
      }
    }
}

/* ------------------------------------------------------------------------------------ */
/* calc_dt_kernel_c.c:31-179.  The per-cell minimum is written to dt_min (work_array1) as the
 * reference does; control/jldt/kldt are the constants the reference returns (:159-163). */
void calc_dt_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *gsmall, double *gbig,
                       double *mindt, double *dtcsafe, double *dtusafe, double *dtvsafe,
                       double *dtdivsafe, double *xarea, double *yarea, double *cellx, double *celly,
                       double *celldx, double *celldy, double *volume, double *density0,
                       double *energy0, double *pressure, double *viscosity, double *soundspeed,
                       double *xvel0, double *yvel0, double *dt_min, double *dtminval,
                       int *dtlcontrol, double *xlpos, double *ylpos, int *jldt, int *kldt, int *smll) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const double g_small = *gsmall, g_big = *gbig;
  const double dtc_safe = *dtcsafe, dtu_safe = *dtusafe, dtv_safe = *dtvsafe, dtdiv_safe = *dtdivsafe;
  double dt_min_val = g_big;
  (void)cellx; (void)celly; (void)smll;
#pragma omp parallel for reduction(min : dt_min_val)
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      const double dsx = celldx[I1X(j)], dsy = celldy[I1Y(k)];
      const double vol = volume[IC(j, k)];
      double cc = soundspeed[IC(j, k)] * soundspeed[IC(j, k)];
      cc = cc + 2.0 * viscosity[IC(j, k)] / density0[IC(j, k)];
      cc = DMAX(sqrt(cc), g_small);
      const double dtct = dtc_safe * DMIN(dsx, dsy) / cc;
      double div = 0.0;
      double dv1 = (xvel0[IV(j, k)] + xvel0[IV(j, k + 1)]) * xarea[IV(j, k)];
      double dv2 = (xvel0[IV(j + 1, k)] + xvel0[IV(j + 1, k + 1)]) * xarea[IV(j + 1, k)];
      div = div + dv2 - dv1;
      const double dtut = dtu_safe * 2.0 * vol / DMAX(fabs(dv1), DMAX(fabs(dv2), g_small * vol));
      dv1 = (yvel0[IV(j, k)] + yvel0[IV(j + 1, k)]) * yarea[IC(j, k)];
      dv2 = (yvel0[IV(j, k + 1)] + yvel0[IV(j + 1, k + 1)]) * yarea[IC(j, k + 1)];
      div = div + dv2 - dv1;
      const double dtvt = dtv_safe * 2.0 * vol / DMAX(fabs(dv1), DMAX(fabs(dv2), g_small * vol));
      div = div / (2.0 * vol);
      const double dtdivt = (div < -g_small) ? dtdiv_safe * (-1.0 / div) : g_big;
      const double m = DMIN(dtct, DMIN(dtut, DMIN(dtvt, dtdivt)));
      dt_min[IV(j, k)] = m;
      if (m < dt_min_val) dt_min_val = m;
    }
  *dtminval = dt_min_val;
  *dtlcontrol = 1;
  *jldt = 1;
  *kldt = 1;
  /* xlpos / ylpos are passed through unchanged by the reference (:161-162) */
  (void)xlpos; (void)ylpos;
  if (dt_min_val < *mindt) {
    printf("Timestep information:\n");
    printf("j, k                 :%i %i \n", 1, 1);
    printf("x, y                 :%f %f \n", *xlpos, *ylpos);
    printf("timestep : %f\n", dt_min_val);
    printf("density, energy, pressure, soundspeed \n");
    printf("%f %f %f %f \n", density0[IC(1, 1)], energy0[IC(1, 1)], pressure[IC(1, 1)],
           soundspeed[IC(1, 1)]);
  }
}

/* ------------------------------------------------------------------------------------ */
/* PdV_kernel_c.c:32-175.  *prdct==0 is the predictor (half step, level-0 velocities only). */
void pdv_kernel_c_(int *prdct, int *xmin, int *xmax, int *ymin, int *ymax, double *dtbyt,
                   double *xarea, double *yarea, double *volume, double *density0, double *density1,
                   double *energy0, double *energy1, double *pressure, double *viscosity,
                   double *xvel0, double *xvel1, double *yvel0, double *yvel1, double *volume_change) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const int predict = (*prdct == 0);
  const double dt = *dtbyt;
  const double *xa = xvel0, *xb = predict ? xvel0 : xvel1;
  const double *ya = yvel0, *yb = predict ? yvel0 : yvel1;
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      double left = xarea[IV(j, k)] * (xa[IV(j, k)] + xa[IV(j, k + 1)] + xb[IV(j, k)] + xb[IV(j, k + 1)]) * 0.25 * dt;
      double right = xarea[IV(j + 1, k)] * (xa[IV(j + 1, k)] + xa[IV(j + 1, k + 1)] + xb[IV(j + 1, k)] + xb[IV(j + 1, k + 1)]) * 0.25 * dt;
      double bottom = yarea[IC(j, k)] * (ya[IV(j, k)] + ya[IV(j + 1, k)] + yb[IV(j, k)] + yb[IV(j + 1, k)]) * 0.25 * dt;
      double top = yarea[IC(j, k + 1)] * (ya[IV(j, k + 1)] + ya[IV(j + 1, k + 1)] + yb[IV(j, k + 1)] + yb[IV(j + 1, k + 1)]) * 0.25 * dt;
      if (predict) { left = left * 0.5; right = right * 0.5; bottom = bottom * 0.5; top = top * 0.5; }
      const double total = right - left + top - bottom;
      const double vol = volume[IC(j, k)];
      const double vc = vol / (vol + total);
      volume_change[IV(j, k)] = vc;
      const double recip = 1.0 / vol;
      const double de = (pressure[IC(j, k)] / density0[IC(j, k)] + viscosity[IC(j, k)] / density0[IC(j, k)]) * total * recip;
      energy1[IC(j, k)] = energy0[IC(j, k)] - de;
      density1[IC(j, k)] = density0[IC(j, k)] * vc;
    }
}

/* revert_kernel_c.c:32-66 */
void revert_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *density0, double *density1,
                      double *energy0, double *energy1) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      density1[IC(j, k)] = density0[IC(j, k)];
      energy1[IC(j, k)] = energy0[IC(j, k)];
    }
}

/* reset_field_kernel_c.c:30-80 */
void reset_field_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *density0,
                           double *density1, double *energy0, double *energy1, double *xvel0,
                           double *xvel1, double *yvel0, double *yvel1) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1 + 1; k++)
    for (int j = g.x0; j <= g.x1 + 1; j++) {
      if (k <= g.y1 && j <= g.x1) {
        density0[IC(j, k)] = density1[IC(j, k)];
        energy0[IC(j, k)] = energy1[IC(j, k)];
      }
      xvel0[IV(j, k)] = xvel1[IV(j, k)];
      yvel0[IV(j, k)] = yvel1[IV(j, k)];
    }
}

/* ------------------------------------------------------------------------------------ */
/* accelerate_kernel_c.c:30-101 */
void accelerate_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *dtbyt, double *xarea,
                          double *yarea, double *volume, double *density0, double *pressure,
                          double *viscosity, double *xvel0, double *yvel0, double *xvel1,
                          double *yvel1) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const double dt = *dtbyt;
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1 + 1; k++)
    for (int j = g.x0; j <= g.x1 + 1; j++) {
      const size_t c11 = IC(j, k), c01 = IC(j - 1, k), c10 = IC(j, k - 1), c00 = IC(j - 1, k - 1);
      const double nodal_mass = (density0[c00] * volume[c00] + density0[c10] * volume[c10] +
                                 density0[c11] * volume[c11] + density0[c01] * volume[c01]) * 0.25;
      const double s = 0.5 * dt / nodal_mass;
      const double xa1 = xarea[IV(j, k)], xa0 = xarea[IV(j, k - 1)];
      const double ya1 = yarea[IC(j, k)], ya0 = yarea[IC(j - 1, k)];
      double xv = xvel0[IV(j, k)] - s * (xa1 * (pressure[c11] - pressure[c01]) + xa0 * (pressure[c10] - pressure[c00]));
      double yv = yvel0[IV(j, k)] - s * (ya1 * (pressure[c11] - pressure[c10]) + ya0 * (pressure[c01] - pressure[c00]));
      xv = xv - s * (xa1 * (viscosity[c11] - viscosity[c01]) + xa0 * (viscosity[c10] - viscosity[c00]));
      yv = yv - s * (ya1 * (viscosity[c11] - viscosity[c10]) + ya0 * (viscosity[c01] - viscosity[c00]));
      xvel1[IV(j, k)] = xv;
      yvel1[IV(j, k)] = yv;
    }
}

/* flux_calc_kernel_c.c:29-77 */
void flux_calc_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *dtbyt, double *xarea,
                         double *yarea, double *xvel0, double *yvel0, double *xvel1, double *yvel1,
                         double *vol_flux_x, double *vol_flux_y) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const double dt = *dtbyt;
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1 + 1; k++)
    for (int j = g.x0; j <= g.x1 + 1; j++) {
      if (k <= g.y1)
        vol_flux_x[IV(j, k)] = 0.25 * dt * xarea[IV(j, k)] *
                               (xvel0[IV(j, k)] + xvel0[IV(j, k + 1)] + xvel1[IV(j, k)] + xvel1[IV(j, k + 1)]);
      if (j <= g.x1)
        vol_flux_y[IC(j, k)] = 0.25 * dt * yarea[IC(j, k)] *
                               (yvel0[IV(j, k)] + yvel0[IV(j + 1, k)] + yvel1[IV(j, k)] + yvel1[IV(j + 1, k)]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* van-Leer limited donor value used by advec_cell for both density and energy
 * (advec_cell_kernel_c.c:124-131,140-147): returns the limiter term. */
static double cell_limiter(double one_minus_sigma, double diffuw, double diffdw, double sigma3,
                           double sigma4) {
  if (diffuw * diffdw > 0.0) {
    const double one_by_six = 1.0 / 6.0;
    const double sgn = (diffdw < 0.0) ? -1.0 : 1.0; /* SIGN(1.0,diffdw) */
    return one_minus_sigma * sgn *
           DMIN(fabs(diffuw), DMIN(fabs(diffdw), one_by_six * (sigma3 * fabs(diffuw) + sigma4 * fabs(diffdw))));
  }
  return 0.0;
}

/* advec_cell_kernel_c.c:30-297.  work arrays (row length nx+5): pre_vol, post_vol, pre_mass,
 * post_mass, advec_vol, post_ener, ener_flux -- written exactly as the reference writes them. */
void advec_cell_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, int *dr, int *swp_nmbr,
                          double *vertexdx, double *vertexdy, double *volume, double *density1,
                          double *energy1, double *mass_flux_x, double *vol_flux_x,
                          double *mass_flux_y, double *vol_flux_y, double *pre_vol, double *post_vol,
                          double *pre_mass, double *post_mass, double *advec_vol, double *post_ener,
                          double *ener_flux) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const int dir = *dr, sweep = *swp_nmbr;
  const int xdir = (dir == 1);
  if (dir != 1 && dir != 2) return;
  /* stage 1: volumes before/after this sweep, over the whole padded range (:72-102, :182-214) */
#pragma omp parallel for
  for (int k = g.y0 - 2; k <= g.y1 + 2; k++)
    for (int j = g.x0 - 2; j <= g.x1 + 2; j++) {
      const double vol = volume[IC(j, k)];
      const double fx0 = vol_flux_x[IV(j, k)], fx1 = vol_flux_x[IV(j + 1, k)];
      const double fy0 = vol_flux_y[IC(j, k)], fy1 = vol_flux_y[IC(j, k + 1)];
      double pre, post;
      if (sweep == 1) {
        if (xdir) { pre = vol + (fx1 - fx0 + fy1 - fy0); post = pre - (fx1 - fx0); }
        else      { pre = vol + (fy1 - fy0 + fx1 - fx0); post = pre - (fy1 - fy0); }
      } else {
        if (xdir) pre = vol + fx1 - fx0; else pre = vol + fy1 - fy0;
        post = vol;
      }
      pre_vol[IV(j, k)] = pre;
      post_vol[IV(j, k)] = post;
    }
  /* stage 2: limited mass and energy fluxes through the faces normal to the sweep (:104-151, :215-262) */
  const int jhi = xdir ? g.x1 + 2 : g.x1, khi = xdir ? g.y1 : g.y1 + 2;
#pragma omp parallel for
  for (int k = g.y0; k <= khi; k++)
    for (int j = g.x0; j <= jhi; j++) {
      const double vf = xdir ? vol_flux_x[IV(j, k)] : vol_flux_y[IC(j, k)];
      const int s = xdir ? j : k;             /* position along the sweep */
      const int smax = (xdir ? g.x1 : g.y1) + 2;
      int up, don, down, dif;
      if (vf > 0.0) { up = s - 2; don = s - 1; down = s; dif = don; }
      else          { up = DMIN(s + 1, smax); don = s; down = s - 1; dif = up; }
#define AT(s_) (xdir ? IC((s_), k) : IC(j, (s_)))
#define ATV(s_) (xdir ? IV((s_), k) : IV(j, (s_)))
      const double *vdx = xdir ? vertexdx : vertexdy;
      const int i1s = xdir ? I1X(s) : I1Y(s), i1d = xdir ? I1X(dif) : I1Y(dif);
      const double sigmat = fabs(vf / pre_vol[ATV(don)]);
      const double sigma3 = (1.0 + sigmat) * (vdx[i1s] / vdx[i1d]);
      const double sigma4 = 2.0 - sigmat;
      double diffuw = density1[AT(don)] - density1[AT(up)];
      double diffdw = density1[AT(down)] - density1[AT(don)];
      double limiter = cell_limiter(1.0 - sigmat, diffuw, diffdw, sigma3, sigma4);
      const double mf = vf * (density1[AT(don)] + limiter);
      if (xdir) mass_flux_x[IV(j, k)] = mf; else mass_flux_y[IC(j, k)] = mf;
      const double sigmam = fabs(mf) / (density1[AT(don)] * pre_vol[ATV(don)]);
      diffuw = energy1[AT(don)] - energy1[AT(up)];
      diffdw = energy1[AT(down)] - energy1[AT(don)];
      limiter = cell_limiter(1.0 - sigmam, diffuw, diffdw, sigma3, sigma4);
      ener_flux[IV(j, k)] = mf * (energy1[AT(don)] + limiter);
#undef AT
#undef ATV
    }
  /* stage 3: conservative cell update (:153-177, :264-290) */
#pragma omp parallel for
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      const size_t w = IV(j, k), wn = xdir ? IV(j + 1, k) : IV(j, k + 1);
      double mf0, mf1, vf0, vf1;
      if (xdir) { mf0 = mass_flux_x[IV(j, k)]; mf1 = mass_flux_x[IV(j + 1, k)]; vf0 = vol_flux_x[IV(j, k)]; vf1 = vol_flux_x[IV(j + 1, k)]; }
      else      { mf0 = mass_flux_y[IC(j, k)]; mf1 = mass_flux_y[IC(j, k + 1)]; vf0 = vol_flux_y[IC(j, k)]; vf1 = vol_flux_y[IC(j, k + 1)]; }
      pre_mass[w] = density1[IC(j, k)] * pre_vol[w];
      post_mass[w] = pre_mass[w] + mf0 - mf1;
      post_ener[w] = (energy1[IC(j, k)] * pre_mass[w] + ener_flux[w] - ener_flux[wn]) / post_mass[w];
      advec_vol[w] = pre_vol[w] + vf0 - vf1;
      density1[IC(j, k)] = post_mass[w] / advec_vol[w];
      energy1[IC(j, k)] = post_ener[w];
    }
}

/* ------------------------------------------------------------------------------------ */
/* advec_mom_kernel_c.c:32-286.  work arrays (row length nx+5): node_flux, node_mass_post,
 * node_mass_pre, mom_flux, pre_vol, post_vol. */
void advec_mom_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *vel1,
                         double *mass_flux_x, double *vol_flux_x, double *mass_flux_y,
                         double *vol_flux_y, double *volume, double *density1, double *node_flux,
                         double *node_mass_post, double *node_mass_pre, double *mom_flux,
                         double *pre_vol, double *post_vol, double *celldx, double *celldy,
                         int *whch_vl, int *swp_nmbr, int *drctn) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const int direction = *drctn;
  const int mom_sweep = direction + 2 * (*swp_nmbr - 1);
  (void)whch_vl; /* the C kernels rebuild the node arrays for both velocity components */
  /* volumes (:69-121) */
  if (mom_sweep >= 1 && mom_sweep <= 4) {
#pragma omp parallel for
    for (int k = g.y0 - 2; k <= g.y1 + 2; k++)
      for (int j = g.x0 - 2; j <= g.x1 + 2; j++) {
        const double vol = volume[IC(j, k)];
        const double fx0 = vol_flux_x[IV(j, k)], fx1 = vol_flux_x[IV(j + 1, k)];
        const double fy0 = vol_flux_y[IC(j, k)], fy1 = vol_flux_y[IC(j, k + 1)];
        double pre, post;
        switch (mom_sweep) {
          case 1: post = vol + fy1 - fy0; pre = post + fx1 - fx0; break;
          case 2: post = vol + fx1 - fx0; pre = post + fy1 - fy0; break;
          case 3: post = vol; pre = post + fy1 - fy0; break;
          default: post = vol; pre = post + fx1 - fx0; break;
        }
        post_vol[IV(j, k)] = post;
        pre_vol[IV(j, k)] = pre;
      }
  }
  if (direction == 1) {
    /* node_flux (:124-134), node_mass_post (:135-150), node_mass_pre (:151-158) */
#pragma omp parallel for
    for (int k = g.y0; k <= g.y1 + 1; k++)
      for (int j = g.x0 - 2; j <= g.x1 + 2; j++)
        node_flux[IV(j, k)] = 0.25 * (mass_flux_x[IV(j, k - 1)] + mass_flux_x[IV(j, k)] +
                                      mass_flux_x[IV(j + 1, k - 1)] + mass_flux_x[IV(j + 1, k)]);
#pragma omp parallel for
    for (int k = g.y0; k <= g.y1 + 1; k++)
      for (int j = g.x0 - 1; j <= g.x1 + 2; j++)
        node_mass_post[IV(j, k)] = 0.25 * (density1[IC(j, k - 1)] * post_vol[IV(j, k - 1)] +
                                           density1[IC(j, k)] * post_vol[IV(j, k)] +
                                           density1[IC(j - 1, k - 1)] * post_vol[IV(j - 1, k - 1)] +
                                           density1[IC(j - 1, k)] * post_vol[IV(j - 1, k)]);
#pragma omp parallel for
    for (int k = g.y0; k <= g.y1 + 1; k++)
      for (int j = g.x0 - 1; j <= g.x1 + 2; j++)
        node_mass_pre[IV(j, k)] = node_mass_post[IV(j, k)] - node_flux[IV(j - 1, k)] + node_flux[IV(j, k)];
    /* limited momentum flux (:159-189) */
#pragma omp parallel for
    for (int k = g.y0; k <= g.y1 + 1; k++)
      for (int j = g.x0 - 1; j <= g.x1 + 1; j++) {
        const double nf = node_flux[IV(j, k)];
        int up, don, down, dif;
        if (nf < 0.0) { up = j + 2; don = j + 1; down = j; dif = don; }
        else          { up = j - 1; don = j; down = j + 1; dif = up; }
        const double sigma = fabs(nf) / node_mass_pre[IV(don, k)];
        const double width = celldx[I1X(j)];
        const double vdiffuw = vel1[IV(don, k)] - vel1[IV(up, k)];
        const double vdiffdw = vel1[IV(down, k)] - vel1[IV(don, k)];
        double limiter = 0.0;
        if (vdiffuw * vdiffdw > 0.0) {
          const double auw = fabs(vdiffuw), adw = fabs(vdiffdw);
          const double wind = (vdiffdw <= 0.0) ? -1.0 : 1.0;
          limiter = wind * DMIN(width * ((2.0 - sigma) * adw / width + (1.0 + sigma) * auw / celldx[I1X(dif)]) / 6.0,
                                DMIN(auw, adw));
        }
        const double advec_vel = vel1[IV(don, k)] + (1.0 - sigma) * limiter;
        mom_flux[IV(j, k)] = advec_vel * nf;
      }
    /* velocity update (:191-201) */
#pragma omp parallel for
    for (int k = g.y0; k <= g.y1 + 1; k++)
      for (int j = g.x0; j <= g.x1 + 1; j++)
        vel1[IV(j, k)] = (vel1[IV(j, k)] * node_mass_pre[IV(j, k)] + mom_flux[IV(j - 1, k)] - mom_flux[IV(j, k)]) /
                         node_mass_post[IV(j, k)];
  } else if (direction == 2) {
    /* (:203-286) */
#pragma omp parallel for
    for (int k = g.y0 - 2; k <= g.y1 + 2; k++)
      for (int j = g.x0; j <= g.x1 + 1; j++)
        node_flux[IV(j, k)] = 0.25 * (mass_flux_y[IC(j - 1, k)] + mass_flux_y[IC(j, k)] +
                                      mass_flux_y[IC(j - 1, k + 1)] + mass_flux_y[IC(j, k + 1)]);
#pragma omp parallel for
    for (int k = g.y0 - 1; k <= g.y1 + 2; k++)
      for (int j = g.x0; j <= g.x1 + 1; j++)
        node_mass_post[IV(j, k)] = 0.25 * (density1[IC(j, k - 1)] * post_vol[IV(j, k - 1)] +
                                           density1[IC(j, k)] * post_vol[IV(j, k)] +
                                           density1[IC(j - 1, k - 1)] * post_vol[IV(j - 1, k - 1)] +
                                           density1[IC(j - 1, k)] * post_vol[IV(j - 1, k)]);
#pragma omp parallel for
    for (int k = g.y0 - 1; k <= g.y1 + 2; k++)
      for (int j = g.x0; j <= g.x1 + 1; j++)
        node_mass_pre[IV(j, k)] = node_mass_post[IV(j, k)] - node_flux[IV(j, k - 1)] + node_flux[IV(j, k)];
#pragma omp parallel for
    for (int k = g.y0 - 1; k <= g.y1 + 1; k++)
      for (int j = g.x0; j <= g.x1 + 1; j++) {
        const double nf = node_flux[IV(j, k)];
        int up, don, down, dif;
        if (nf < 0.0) { up = k + 2; don = k + 1; down = k; dif = don; }
        else          { up = k - 1; don = k; down = k + 1; dif = up; }
        const double sigma = fabs(nf) / node_mass_pre[IV(j, don)];
        const double width = celldy[I1Y(k)];
        const double vdiffuw = vel1[IV(j, don)] - vel1[IV(j, up)];
        const double vdiffdw = vel1[IV(j, down)] - vel1[IV(j, don)];
        double limiter = 0.0;
        if (vdiffuw * vdiffdw > 0.0) {
          const double auw = fabs(vdiffuw), adw = fabs(vdiffdw);
          const double wind = (vdiffdw <= 0.0) ? -1.0 : 1.0;
          limiter = wind * DMIN(width * ((2.0 - sigma) * adw / width + (1.0 + sigma) * auw / celldy[I1Y(dif)]) / 6.0,
                                DMIN(auw, adw));
        }
        const double advec_vel = vel1[IV(j, don)] + (1.0 - sigma) * limiter;
        mom_flux[IV(j, k)] = advec_vel * nf;
      }
#pragma omp parallel for
    for (int k = g.y0; k <= g.y1 + 1; k++)
      for (int j = g.x0; j <= g.x1 + 1; j++)
        vel1[IV(j, k)] = (vel1[IV(j, k)] * node_mass_pre[IV(j, k)] + mom_flux[IV(j, k - 1)] - mom_flux[IV(j, k)]) /
                         node_mass_post[IV(j, k)];
  }
}

/* ------------------------------------------------------------------------------------ */
/* field_summary_kernel_c.c:30-98.  Serial row-major accumulation (the order a 1-thread run of
 * the reference uses); an OpenMP build of the reference differs in the last digits. */
void field_summary_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *volume,
                             double *density0, double *energy0, double *pressure, double *xvel0,
                             double *yvel0, double *vl, double *mss, double *ien, double *ken,
                             double *prss) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  double vol = 0.0, mass = 0.0, ie = 0.0, ke = 0.0, press = 0.0;
  for (int k = g.y0; k <= g.y1; k++)
    for (int j = g.x0; j <= g.x1; j++) {
      double vsqrd = 0.0;
      for (int kv = k; kv <= k + 1; kv++)
        for (int jv = j; jv <= j + 1; jv++)
          vsqrd = vsqrd + 0.25 * (xvel0[IV(jv, kv)] * xvel0[IV(jv, kv)] + yvel0[IV(jv, kv)] * yvel0[IV(jv, kv)]);
      const double cell_vol = volume[IC(j, k)];
      const double cell_mass = cell_vol * density0[IC(j, k)];
      vol = vol + cell_vol;
      mass = mass + cell_mass;
      ie = ie + cell_mass * energy0[IC(j, k)];
      ke = ke + cell_mass * 0.5 * vsqrd;
      press = press + cell_vol * pressure[IC(j, k)];
    }
  *vl = vol; *mss = mass; *ien = ie; *ken = ke; *prss = press;
}

/* ------------------------------------------------------------------------------------ */
/* update_halo_kernel_c.c:32-716 -- reflective boundary on external faces.
 * Per field type (x_inc,y_inc) and "m" (0 for cell-centred data, 1 otherwise) the 60 loops of the
 * reference reduce to:
 *   bottom: f(j,1-k)        = sy * f(j, m+k)              j in x_min-d .. x_max+x_inc+d
 *   top:    f(j,ny+y_inc+k) = sy * f(j, ny+y_inc+(1-m)-k)
 *   left:   f(1-j,k)        = sx * f(m+j, k)              k in y_min-d .. y_max+y_inc+d
 *   right:  f(nx+x_inc+j,k) = sx * f(nx+x_inc+(1-m)-j, k)
 * with the literal lower bound 1 the reference hard-codes.  Order per field: bottom, top, left, right. */
static void reflect_field(const grid_t g, const int *ext, double *f, int x_inc, int y_inc, int m,
                          double sx, double sy, int depth) {
  const int row = g.x1 + 4 + x_inc;
  const int nx = g.x1, ny = g.y1;
#define F(j, k) f[(size_t)((k) - (g.y0 - 2)) * row + ((j) - (g.x0 - 2))]
  if (ext[2])
    for (int j = g.x0 - depth; j <= g.x1 + x_inc + depth; j++)
      for (int k = 1; k <= depth; k++) F(j, 1 - k) = sy * F(j, m + k);
  if (ext[3])
    for (int j = g.x0 - depth; j <= g.x1 + x_inc + depth; j++)
      for (int k = 1; k <= depth; k++) F(j, ny + y_inc + k) = sy * F(j, ny + y_inc + (1 - m) - k);
  if (ext[0])
    for (int k = g.y0 - depth; k <= g.y1 + y_inc + depth; k++)
      for (int j = 1; j <= depth; j++) F(1 - j, k) = sx * F(m + j, k);
  if (ext[1])
    for (int k = g.y0 - depth; k <= g.y1 + y_inc + depth; k++)
      for (int j = 1; j <= depth; j++) F(nx + x_inc + j, k) = sx * F(nx + x_inc + (1 - m) - j, k);
#undef F
}

void update_halo_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, int *chunk_neighbours,
                           int *tile_neighbours, double *density0, double *energy0, double *pressure,
                           double *viscosity, double *soundspeed, double *density1, double *energy1,
                           double *xvel0, double *yvel0, double *xvel1, double *yvel1,
                           double *vol_flux_x, double *vol_flux_y, double *mass_flux_x,
                           double *mass_flux_y, int *fields, int *dpth) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const int depth = *dpth;
  int ext[4]; /* left, right, bottom, top */
  for (int f = 0; f < 4; f++) ext[f] = (chunk_neighbours[f] == -1 && tile_neighbours[f] == -1);
  /* field ids: data.f90:51-66 (1-based) */
  if (fields[0] == 1) reflect_field(g, ext, density0, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[1] == 1) reflect_field(g, ext, density1, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[2] == 1) reflect_field(g, ext, energy0, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[3] == 1) reflect_field(g, ext, energy1, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[4] == 1) reflect_field(g, ext, pressure, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[5] == 1) reflect_field(g, ext, viscosity, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[6] == 1) reflect_field(g, ext, soundspeed, 0, 0, 0, 1.0, 1.0, depth);
  if (fields[7] == 1) reflect_field(g, ext, xvel0, 1, 1, 1, -1.0, 1.0, depth);
  if (fields[8] == 1) reflect_field(g, ext, xvel1, 1, 1, 1, -1.0, 1.0, depth);
  if (fields[9] == 1) reflect_field(g, ext, yvel0, 1, 1, 1, 1.0, -1.0, depth);
  if (fields[10] == 1) reflect_field(g, ext, yvel1, 1, 1, 1, 1.0, -1.0, depth);
  if (fields[11] == 1) reflect_field(g, ext, vol_flux_x, 1, 0, 1, -1.0, 1.0, depth);
  if (fields[13] == 1) reflect_field(g, ext, mass_flux_x, 1, 0, 1, -1.0, 1.0, depth);
  if (fields[12] == 1) reflect_field(g, ext, vol_flux_y, 0, 1, 1, 1.0, -1.0, depth);
  if (fields[14] == 1) reflect_field(g, ext, mass_flux_y, 0, 1, 1, 1.0, -1.0, depth);
}

/* ------------------------------------------------------------------------------------ */
/* pack_kernel_c.c:29-439.  face: 0 left, 1 right, 2 bottom, 3 top; unpack!=0 reverses the copy. */
static void halo_message(int face, int unpack, int *xmin, int *xmax, int *ymin, int *ymax,
                         double *field, double *buffer, int *cell, int *vertex, int *xface,
                         int *yface, int *dpth, int *fld_typ, int *bffr_ffst) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const int depth = *dpth, type = *fld_typ, off = *bffr_ffst;
  int x_inc = 0, y_inc = 0;
  if (type == *cell) { x_inc = 0; y_inc = 0; }
  if (type == *vertex) { x_inc = 1; y_inc = 1; }
  if (type == *xface) { x_inc = 1; y_inc = 0; }
  if (type == *yface) { x_inc = 0; y_inc = 1; }
  const int row = g.x1 + 4 + x_inc;
#define F(j, k) field[(size_t)((k) - (g.y0 - 2)) * row + ((j) - (g.x0 - 2))]
  if (face < 2) {
    for (int k = g.y0 - depth; k <= g.y1 + y_inc + depth; k++)
      for (int j = 1; j <= depth; j++) {
        const int index = off + j + (k + depth - 1) * depth - 1; /* 0-based */
        int src;
        if (face == 0) src = unpack ? g.x0 - j : g.x0 + x_inc - 1 + j;
        else           src = unpack ? g.x1 + x_inc + j : g.x1 + 1 - j;
        if (unpack) F(src, k) = buffer[index]; else buffer[index] = F(src, k);
      }
  } else {
    for (int k = 1; k <= depth; k++)
      for (int j = g.x0 - depth; j <= g.x1 + x_inc + depth; j++) {
        const int index = off + k + (j + depth - 1) * depth - 1;
        int src;
        if (face == 2) src = unpack ? g.y0 - k : g.y0 + y_inc - 1 + k;
        else           src = unpack ? g.y1 + y_inc + k : g.y1 + 1 - k;
        if (unpack) F(j, src) = buffer[index]; else buffer[index] = F(j, src);
      }
  }
#undef F
}
#define PACK_ENTRY(name, face, unpack)                                                            \
  void name(int *xmin, int *xmax, int *ymin, int *ymax, double *field, double *buffer, int *c,    \
            int *v, int *xf, int *yf, int *dpth, int *fld_typ, int *bffr_ffst) {                  \
    halo_message(face, unpack, xmin, xmax, ymin, ymax, field, buffer, c, v, xf, yf, dpth, fld_typ, \
                 bffr_ffst);                                                                      \
  }
PACK_ENTRY(clover_pack_message_left_c_, 0, 0)
PACK_ENTRY(clover_unpack_message_left_c_, 0, 1)
PACK_ENTRY(clover_pack_message_right_c_, 1, 0)
PACK_ENTRY(clover_unpack_message_right_c_, 1, 1)
PACK_ENTRY(clover_pack_message_bottom_c_, 2, 0)
PACK_ENTRY(clover_unpack_message_bottom_c_, 2, 1)
PACK_ENTRY(clover_pack_message_top_c_, 3, 0)
PACK_ENTRY(clover_unpack_message_top_c_, 3, 1)

/* ------------------------------------------------------------------------------------ */
/* initialise_chunk_kernel_c.c:29-125.  xarea(x_max+3,:) and yarea(:,y_max+3) are left untouched
 * (the loops stop at +2), as in the reference. */
void initialise_chunk_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *minx,
                                double *miny, double *dx, double *dy, double *vertexx,
                                double *vertexdx, double *vertexy, double *vertexdy, double *cellx,
                                double *celldx, double *celly, double *celldy, double *volume,
                                double *xarea, double *yarea) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  const double min_x = *minx, min_y = *miny, d_x = *dx, d_y = *dy;
  for (int j = g.x0 - 2; j <= g.x1 + 3; j++) {
    vertexx[I1X(j)] = min_x + d_x * (double)(j - g.x0);
    vertexdx[I1X(j)] = d_x;
  }
  for (int k = g.y0 - 2; k <= g.y1 + 3; k++) {
    vertexy[I1Y(k)] = min_y + d_y * (double)(k - g.y0);
    vertexdy[I1Y(k)] = d_y;
  }
  for (int j = g.x0 - 2; j <= g.x1 + 2; j++) {
    cellx[I1X(j)] = 0.5 * (vertexx[I1X(j)] + vertexx[I1X(j + 1)]);
    celldx[I1X(j)] = d_x;
  }
  for (int k = g.y0 - 2; k <= g.y1 + 2; k++) {
    celly[I1Y(k)] = 0.5 * (vertexy[I1Y(k)] + vertexy[I1Y(k + 1)]);
    celldy[I1Y(k)] = d_y;
  }
  for (int k = g.y0 - 2; k <= g.y1 + 2; k++)
    for (int j = g.x0 - 2; j <= g.x1 + 2; j++) {
      volume[IC(j, k)] = d_x * d_y;
      xarea[IV(j, k)] = celldy[I1Y(k)];
      yarea[IC(j, k)] = celldx[I1X(j)];
    }
}

/* generate_chunk_kernel_c.c:33-162.  Quirks kept: circle and point geometries set energy0 to the
 * state DENSITY (:134,145); the point test reads vertexy at index j (:143). */
void generate_chunk_kernel_c_(int *xmin, int *xmax, int *ymin, int *ymax, double *vertexx,
                              double *vertexy, double *cellx, double *celly, double *density0,
                              double *energy0, double *xvel0, double *yvel0, int *nmbr_f_stts,
                              double *state_density, double *state_energy, double *state_xvel,
                              double *state_yvel, double *state_xmin, double *state_xmax,
                              double *state_ymin, double *state_ymax, double *state_radius,
                              int *state_geometry, int *g_rct, int *g_crc, int *g_pnt) {
  const grid_t g = make_grid(xmin, xmax, ymin, ymax);
  for (int k = g.y0 - 2; k <= g.y1 + 2; k++)
    for (int j = g.x0 - 2; j <= g.x1 + 2; j++) {
      energy0[IC(j, k)] = state_energy[0];
      density0[IC(j, k)] = state_density[0];
      xvel0[IV(j, k)] = state_xvel[0];
      yvel0[IV(j, k)] = state_yvel[0];
    }
  for (int s = 1; s < *nmbr_f_stts; s++) {
    const double x_cent = state_xmin[s], y_cent = state_ymin[s];
    for (int k = g.y0 - 2; k <= g.y1 + 2; k++)
      for (int j = g.x0 - 2; j <= g.x1 + 2; j++) {
        int hit = 0;
        double e = state_energy[s];
        if (state_geometry[s] == *g_rct) {
          hit = vertexx[I1X(j + 1)] >= state_xmin[s] && vertexx[I1X(j)] < state_xmax[s] &&
                vertexy[I1Y(k + 1)] >= state_ymin[s] && vertexy[I1Y(k)] < state_ymax[s];
        } else if (state_geometry[s] == *g_crc) {
          const double radius = sqrt((cellx[I1X(j)] - x_cent) * (cellx[I1X(j)] - x_cent) +
                                     (celly[I1Y(k)] - y_cent) * (celly[I1Y(k)] - y_cent));
          hit = radius <= state_radius[s];
          e = state_density[s];
        } else if (state_geometry[s] == *g_pnt) {
          hit = vertexx[I1X(j)] == x_cent && vertexy[I1X(j)] == y_cent;
          e = state_density[s];
        }
        if (hit) {
          density0[IC(j, k)] = state_density[s];
          energy0[IC(j, k)] = e;
          for (int kt = k; kt <= k + 1; kt++)
            for (int jt = j; jt <= j + 1; jt++) {
              xvel0[IV(jt, kt)] = state_xvel[s];
              yvel0[IV(jt, kt)] = state_yvel[s];
            }
        }
      }
  }
}

/* timer_c.c:32-38 */
void timer_c_(double *elapsed_time) {
  struct timeval t;
  gettimeofday(&t, NULL);
  *elapsed_time = t.tv_sec + t.tv_usec * 1.0E-6;
}
