/*
 * ref_shim.cu -- build recipe glue for oracle/_ref/libref_v9.so.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the UNMODIFIED reference translation unit where it lies (include path set by
 * oracle/Makefile to /root/reference; no reference source is copied into this repo) and
 * exposes two of its functions with C linkage:
 *   ref_cpu_attention  -> cpu_attention               (flash_attention.cu:668-697)
 *   ref_v9_dispatch    -> flash_attention_v9_dispatch (flash_attention.cu:606-663)
 * The reference's main() (flash_attention.cu:702) is renamed so the TU can live in a
 * shared object.  Flags follow the reference Makefile:4 (-O3 --use_fast_math) with the
 * arch switched to sm_100a ("V9 rebuilt for the box").
 */
#define main ref_harness_main
#include "flash_attention.cu"
#undef main

extern "C" void ref_cpu_attention(const void* Q, const void* K, const void* V, void* O,
                                  int B, int H, int N, int D, int causal)
{
    cpu_attention((const half*)Q, (const half*)K, (const half*)V, (half*)O, B, H, N, D, causal != 0);
}

extern "C" void ref_v9_dispatch(const void* Q, const void* K, const void* V, void* O,
                                int B, int H, int N, int D, int causal, void* stream)
{
    flash_attention_v9_dispatch((const half*)Q, (const half*)K, (const half*)V, (half*)O,
                                nullptr, nullptr, B, H, N, D, causal != 0, (cudaStream_t)stream);
}

extern "C" int ref_harness(void) { return ref_harness_main(); }
