// Template dispatch for the fine-grid kernels; instantiated once per storage precision in
// dslash_d.cu / dslash_s.cu / dslash_h.cu so the three precisions compile in parallel.
#pragma once
#include "dslash.cuh"
#include "dslash_api.h"

namespace qb {

template <typename Store, int RECON, bool GHOST>
static void launch_dslash_recon(const DslashParam &p, bool twist_in, bool has_x, int block, cudaStream_t s, int clover = 0) {
  if (clover) {  // hop + inverse twisted-clover block in one launch (single fields, no input twist)
    if (p.nbatch > 1 || twist_in) QB_ERROR("fused clover hop: single fields without input twist only");
    const int nbc = div_up(p.site_count, block);
    if (nbc == 0) return;
    if (clover == 1) {
      if (has_x) dslash_kernel<Store, RECON, false, true, GHOST, false, 1><<<nbc, block, 0, s>>>(p);
      else dslash_kernel<Store, RECON, false, false, GHOST, false, 1><<<nbc, block, 0, s>>>(p);
    } else {
      if (has_x) dslash_kernel<Store, RECON, false, true, GHOST, false, 2><<<nbc, block, 0, s>>>(p);
      else dslash_kernel<Store, RECON, false, false, GHOST, false, 2><<<nbc, block, 0, s>>>(p);
    }
    QB_CHECK_LAUNCH();
    return;
  }
  if (p.nbatch > 1) {  // batched fields: 32 sites x nbatch members per CTA
    if (GHOST || Store::scaled) QB_ERROR("batched hop: unpartitioned lattices and fp32 / fp64 fields only");
    if (p.nbatch > DSLASH_BATCH_MAX) QB_ERROR("batched hop: at most %d members per launch", DSLASH_BATCH_MAX);
    const int nbb = div_up(p.site_count, 32);
    const dim3 bd(32, p.nbatch);
    if (nbb == 0) return;
    if constexpr (!GHOST && !Store::scaled) {
      if (twist_in) {
        if (has_x) dslash_kernel<Store, RECON, true, true, false, true><<<nbb, bd, 0, s>>>(p);
        else dslash_kernel<Store, RECON, true, false, false, true><<<nbb, bd, 0, s>>>(p);
      } else {
        if (has_x) dslash_kernel<Store, RECON, false, true, false, true><<<nbb, bd, 0, s>>>(p);
        else dslash_kernel<Store, RECON, false, false, false, true><<<nbb, bd, 0, s>>>(p);
      }
    }
    QB_CHECK_LAUNCH();
    return;
  }
  const int nb = div_up(p.site_count, block);
  if (nb == 0) return;
  if (twist_in) {
    if (has_x) dslash_kernel<Store, RECON, true, true, GHOST><<<nb, block, 0, s>>>(p);
    else dslash_kernel<Store, RECON, true, false, GHOST><<<nb, block, 0, s>>>(p);
  } else {
    if (has_x) dslash_kernel<Store, RECON, false, true, GHOST><<<nb, block, 0, s>>>(p);
    else dslash_kernel<Store, RECON, false, false, GHOST><<<nb, block, 0, s>>>(p);
  }
  QB_CHECK_LAUNCH();
}

template <typename Store>
void launch_dslash_T(const DslashParam &p, int recon, bool twist_in, bool has_x, bool ghost, int block, cudaStream_t s, int clover) {
  // ghost = false: no site of this launch has a neighbour in a ghost zone (unpartitioned lattice, interior launch)
  if (ghost) {
    if (recon == 18) launch_dslash_recon<Store, 18, true>(p, twist_in, has_x, block, s, clover);
    else if (recon == 12) launch_dslash_recon<Store, 12, true>(p, twist_in, has_x, block, s, clover);
    else if (recon == 8) launch_dslash_recon<Store, 8, true>(p, twist_in, has_x, block, s, clover);
    else QB_ERROR("unsupported reconstruct %d", recon);
  } else {
    if (recon == 18) launch_dslash_recon<Store, 18, false>(p, twist_in, has_x, block, s, clover);
    else if (recon == 12) launch_dslash_recon<Store, 12, false>(p, twist_in, has_x, block, s, clover);
    else if (recon == 8) launch_dslash_recon<Store, 8, false>(p, twist_in, has_x, block, s, clover);
    else QB_ERROR("unsupported reconstruct %d", recon);
  }
}

template <typename Store>
void launch_pack_T(const PackParam &p, bool twist_in, cudaStream_t s) {
  const int n = p.thread_off[4];
  if (n == 0) return;
  if (twist_in) pack_kernel<Store, true><<<div_up(n, 128), 128, 0, s>>>(p);
  else pack_kernel<Store, false><<<div_up(n, 128), 128, 0, s>>>(p);
  QB_CHECK_LAUNCH();
}

template <typename Store>
void launch_twist_T(void *out, float *out_norm, const void *in, const float *in_norm, long stride, int n, double pr, double qr, cudaStream_t s) {
  twist_kernel<Store><<<div_up(n, 256), 256, 0, s>>>(out, out_norm, in, in_norm, stride, n, pr, qr);
  QB_CHECK_LAUNCH();
}

#define QB_INSTANTIATE_DSLASH(Store)                                                                          \
  template void launch_dslash_T<Store>(const DslashParam &, int, bool, bool, bool, int, cudaStream_t, int);              \
  template void launch_pack_T<Store>(const PackParam &, bool, cudaStream_t);                                  \
  template void launch_twist_T<Store>(void *, float *, const void *, const float *, long, int, double, double, cudaStream_t);

}  // namespace qb
