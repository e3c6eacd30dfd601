/* TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libmgref.so, see mg_ref_shim.cpp).
 * Run-time type information of the reference's DEVICE field classes.  The reference's host code dynamic_casts to them (always failing
 * here: no device field is ever created), so the type_info objects must exist although the classes themselves (cuda_*_field.cu) are
 * not compiled into the CPU-only oracle.  Itanium C++ ABI layout of a single-inheritance class type_info: { vptr, name, base }.
 * Plain C so that the mangled names can be spelled out. */
extern void *_ZTVN10__cxxabiv120__si_class_type_infoE[];
extern char _ZTIN4quda16ColorSpinorFieldE[];   /* typeinfo for quda::ColorSpinorField (lib/color_spinor_field.cpp) */
extern char _ZTIN4quda10GaugeFieldE[];         /* typeinfo for quda::GaugeField (lib/gauge_field.cpp) */
struct mgref_type_info { const void *vptr; const char *name; const void *base; };
struct mgref_type_info _ZTIN4quda20cudaColorSpinorFieldE = {&_ZTVN10__cxxabiv120__si_class_type_infoE[2], "N4quda20cudaColorSpinorFieldE", _ZTIN4quda16ColorSpinorFieldE};
struct mgref_type_info _ZTIN4quda14cudaGaugeFieldE = {&_ZTVN10__cxxabiv120__si_class_type_infoE[2], "N4quda14cudaGaugeFieldE", _ZTIN4quda10GaugeFieldE};

/* Entry points of the parts of the reference that are NOT built into the CPU-only oracle (device fields, generic copy kernels, clover
 * construction, the CUDA driver API).  The host paths driven by mg_ref_shim.cpp never reach them; python's dlopen binds every symbol at
 * load time, so they have to exist.  Each one aborts with its name if it is ever called. */
#include <stdio.h>
#include <stdlib.h>
static void mgref_unavailable(const char *what) {
  fprintf(stderr, "oracle/_ref/libmgref.so: %s is not part of the CPU-only reference build\n", what);
  abort();
}
#define MGREF_STUB(name) void name(void) { mgref_unavailable(#name); }
MGREF_STUB(_ZN4quda11qudaMemcpy_EPvPKvm14cudaMemcpyKindPKcS5_i)
MGREF_STUB(_ZN4quda13computeCloverERNS_11CloverFieldERKNS_10GaugeFieldEd19QudaFieldLocation_s)
MGREF_STUB(_ZN4quda13genericSourceERNS_19cpuColorSpinorFieldE16QudaSourceType_siii)
MGREF_STUB(_ZN4quda14genericCompareERKNS_19cpuColorSpinorFieldES2_i)
MGREF_STUB(_ZN4quda15applyGaugePhaseERNS_10GaugeFieldE)
MGREF_STUB(_ZN4quda16copyGenericGaugeERNS_10GaugeFieldERKS0_19QudaFieldLocation_sPvS5_PS5_S6_i)
MGREF_STUB(_ZN4quda17copyGenericCloverERNS_11CloverFieldERKS0_b19QudaFieldLocation_sPvS5_S5_S5_)
MGREF_STUB(_ZN4quda18genericPrintVectorERNS_19cpuColorSpinorFieldEj)
MGREF_STUB(_ZN4quda20cudaColorSpinorFieldC1ERKNS_16ColorSpinorFieldERKNS_16ColorSpinorParamE)
MGREF_STUB(_ZN4quda20cudaColorSpinorFieldC1ERKNS_16ColorSpinorParamE)
MGREF_STUB(_ZN4quda22copyGenericColorSpinorERNS_16ColorSpinorFieldERKS0_19QudaFieldLocation_sPvS5_S5_S5_)
MGREF_STUB(_ZN4quda25extractExtendedGaugeGhostERKNS_10GaugeFieldEiPKiPPvb)
MGREF_STUB(_ZN4quda8maxGaugeERKNS_10GaugeFieldE)
MGREF_STUB(_ZNK4quda14cudaGaugeField12saveCPUFieldERNS_13cpuGaugeFieldE)
MGREF_STUB(_ZNK4quda20cudaColorSpinorField15saveSpinorFieldERNS_16ColorSpinorFieldE)
MGREF_STUB(cuMemAlloc_v2)
MGREF_STUB(cuMemFree_v2)

/* lib/malloc.cpp:pinned_malloc_ page-locks host buffers with cudaHostRegister (the empty cpuCloverField that calculateY takes does so,
 * lib/clover_field.cpp:38).  The oracle also runs where there is no GPU, so inside THIS library the two calls are no-ops: the memory is
 * ordinary aligned host memory and no device ever touches it.  (Defined here, they take precedence over libcudart's for calls made from
 * libmgref.so only.) */
#include <stddef.h>
int cudaHostRegister(void *ptr, size_t size, unsigned int flags) { (void)ptr; (void)size; (void)flags; return 0; }
int cudaHostUnregister(void *ptr) { (void)ptr; return 0; }
