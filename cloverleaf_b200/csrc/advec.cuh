// advec.cuh -- per-face arithmetic of the advective remap shared by advec.cu (register/shuffle kernels) and
// advec_tma.cu (TMA tile kernels).  Every expression keeps the reference's evaluation order.
#pragma once
#include "common.cuh"

namespace clv {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// advec_cell_kernel_c.c:124-131 / :140-147 (and the y twins): the van Leer limited slope
__device__ __forceinline__ double cell_limiter(double one_minus_sigma, double diffuw, double diffdw,
                                               double sigma3, double sigma4) {
  double lim = 0.0;
  if (diffuw * diffdw > 0.0) {
    const double auw = fabs(diffuw), adw = fabs(diffdw);
    const double sgn = (diffdw < 0.0) ? -1.0 : 1.0;
    lim = one_minus_sigma * sgn * dmin(auw, dmin(adw, (1.0 / 6.0) * (sigma3 * auw + sigma4 * adw)));
  }
  return lim;
}

// advec_cell_kernel_c.c:107-151: mass and energy flux through one face, given the face's volume flux,
// the donor cell's pre-volume, the (upwind, donor, downwind) densities/energies and the two widths.
__device__ __forceinline__ void cell_face_flux(double vf, double pre_vol_donor, double d_up, double d_don,
                                               double d_down, double e_up, double e_don, double e_down,
                                               double vd_face, double vd_dif, double& mass_flux,
                                               double& ener_flux) {
  const double sigmat = fabs(ddiv(vf, pre_vol_donor));
  const double sigma3 = (1.0 + sigmat) * (vd_face / vd_dif);
  const double sigma4 = 2.0 - sigmat;
  double limiter = cell_limiter(1.0 - sigmat, d_don - d_up, d_down - d_don, sigma3, sigma4);
  mass_flux = vf * (d_don + limiter);
  const double sigmam = ddiv(fabs(mass_flux), d_don * pre_vol_donor);
  limiter = cell_limiter(1.0 - sigmam, e_don - e_up, e_down - e_don, sigma3, sigma4);
  ener_flux = mass_flux * (e_don + limiter);
}

// advec_mom_kernel_c.c:160-188: limited momentum flux through one node "face"
__device__ __forceinline__ double mom_face_flux(double nf, double node_mass_pre_donor, double v_up,
                                                double v_don, double v_down, double width,
                                                double width_dif) {
  const double sigma = ddiv(fabs(nf), node_mass_pre_donor);
  const double vdiffuw = v_don - v_up;
  const double vdiffdw = v_down - v_don;
  double limiter = 0.0;
  if (vdiffuw * vdiffdw > 0.0) {
    const double auw = fabs(vdiffuw), adw = fabs(vdiffdw);
    const double wind = (vdiffdw <= 0.0) ? -1.0 : 1.0;
    limiter = wind * dmin(width * ((2.0 - sigma) * adw / width + (1.0 + sigma) * auw / width_dif) / 6.0,
                          dmin(auw, adw));
  }
  const double advec_vel = v_don + (1.0 - sigma) * limiter;
  return advec_vel * nf;
}

}  // namespace clv
