/* TEST INFRASTRUCTURE ONLY.  Force-included when the reference's lib/dslash_coarse.cu is compiled for the CPU-only oracle
 * (oracle/Makefile): CUDA 12 removed the un-synchronised __shfl_down the file's DEVICE code uses (lib/dslash_coarse.cu:255).  The oracle
 * runs the file's HOST path only (coarseDslash on the CPU, :263-290); this overload exists so that the unmodified source compiles. */
#ifdef __CUDACC__
template <typename T> __device__ inline T __shfl_down(const T &v, int d) {
  T r;
  const unsigned *s = reinterpret_cast<const unsigned *>(&v);
  unsigned *o = reinterpret_cast<unsigned *>(&r);
  for (unsigned i = 0; i < sizeof(T) / 4; i++) o[i] = __shfl_down_sync(0xffffffffu, s[i], d);
  return r;
}
#endif
