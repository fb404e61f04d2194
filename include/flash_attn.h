/*
 * flash_attn.h -- C ABI of libflashattn_b200.so: FlashAttention forward for NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary for the reference's host launcher
 *     void flash_attention_v9_dispatch(const half* Q, const half* K, const half* V, half* Output,
 *                                      float* splitk_buf_O, float* splitk_buf_ml,
 *                                      int batch_size, int num_heads, int seq_len, int head_dim,
 *                                      bool causal, cudaStream_t stream = 0)
 * (reference flash_attention.cu:606-663).  Same data contract:
 *   - device pointers, FP16, contiguous [B, H, N, D] (row stride D, head stride N*D; FA.cu:119-122)
 *   - softmax scale fixed to 1/sqrt(D) (FA.cu:612)
 *   - causal = lower-triangular including the diagonal: query i sees keys 0..i (FA.cu:250-255, 679)
 *   - O is fully overwritten for rows < N; caller owns every buffer; the call only enqueues work on
 *     `stream` and returns (FA.cu:634-662)
 * Differences, all widening: D may be 64 or 128 (the reference hard-codes 128, FA.cu:613), 64-bit
 * indexing (the reference indexes with int, FA.cu:119-122), errors are returned instead of
 * exit(EXIT_FAILURE) (FA.cu:22-30).  Plain pointers and sizes only; `stream` is a cudaStream_t
 * passed as void* so the header needs no CUDA include.
 */
#ifndef FLASH_ATTN_B200_H
#define FLASH_ATTN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Return codes: 0 = success; > 0 = a cudaError_t from the runtime; < 0 = argument errors below. */
#define FA_OK 0
#define FA_ERR_BAD_HEAD_DIM (-1) /* D not in {64, 128} */
#define FA_ERR_NULL_PTR (-2)
#define FA_ERR_MISALIGNED (-3) /* base pointer not 16-byte aligned (TMA global-address rule) */
#define FA_ERR_BAD_SHAPE (-4)  /* B, H or N < 1, or B*H / N beyond the tensor-map limits */
#define FA_ERR_UNSUPPORTED_ARCH (-5) /* current device is not compute capability 10.x */
#define FA_ERR_TENSORMAP (-6)        /* cuTensorMapEncodeTiled failed */
#define FA_ERR_WORKSPACE (-7)        /* workspace too small / missing for an _ex call; more than 1024 launches recorded into CUDA graphs */
#define FA_ERR_WATCHDOG (-8)         /* an EARLIER kernel of this process timed out on an internal barrier (see flash_attn_status) */

/* Replaces flash_attention_v9_dispatch (FA.cu:606-663).  Argument order is the one the
 * north-star names: q, k, v, o, B, H, N, D, causal, stream. */
int flash_attn_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D,
                   int causal, void* stream);

/* Same operation on BF16 tensors (the reference is FP16 only): Q, K, V and O are bfloat16, the
 * tensor cores take BF16 operands (P is rounded to BF16 as well), S and O accumulate in FP32. */
int flash_attn_fwd_bf16(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D,
                        int causal, void* stream);

/* Extended entry: one K/V block of a longer sequence, for ring context
 * parallelism and split-KV.  Computes attention of the local queries q[B,H,Nq,D] against
 * k/v[B,H,Nkv,D] where the queries sit at global positions q_offset.. and the keys at
 * kv_offset.. (the causal mask compares global positions), and emits the reference's split-K
 * partial format (FA.cu:460-496): o_partial fp32 un-normalised [B*H*Nq, D] and ml [B*H*Nq, 2] =
 * (row max in the scaled-score domain, row sum).  When `accumulate` is non-zero the partials
 * already in o_partial/ml are merged in with the algebra of FA.cu:575-597, so a ring of P hops
 * is P calls on the same buffers.  flash_attn_finalize() then writes O = o_partial / l as FP16. */
int flash_attn_fwd_ex(const void* q, const void* k, const void* v, float* o_partial, float* ml,
                      int B, int H, int Nq, int Nkv, int D, int causal, long long q_offset,
                      long long kv_offset, int accumulate, void* stream);
int flash_attn_finalize(const float* o_partial, const float* ml, void* o, long long rows, int D,
                        void* stream);

/* Context parallelism without partial states (the reference has no multi-GPU code; its split-K partials, FA.cu:460-496,
 * are what the entry above serves).  The queries q[B,H,Nq,D] sit at global positions q_offset.. of a sequence whose visible
 * keys have been GATHERED into one buffer per head: k, v point at B*H heads of `kv_head_rows` rows each, of which the first
 * Nkv are used; key row c is visible to query row r iff c <= r + q_offset (causal) -- keys entirely in the past may stand in
 * any order, the chunk that contains the diagonal comes last.  O[B,H,Nq,D] is written as FP16 like flash_attn_fwd: the
 * accumulators stay in tensor memory across the whole gathered sequence, nothing is merged afterwards.
 * `ready` (optional): int flags, one per `ready_rows` rows of the gathered sequence (a multiple of 128).  The kernel loads
 * rows [i*ready_rows, (i+1)*ready_rows) only after ready[i] != 0, so it may be launched while copy engines are still
 * filling the buffer (flash_attn_peer_copy_2d + flash_attn_stream_write_flag on another stream). */
int flash_attn_fwd_gathered(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nkv, int D,
                            int causal, long long q_offset, long long kv_head_rows, const int* ready, int ready_rows,
                            void* stream);
/* *flag = value as a stream memory operation (cuStreamWriteValue32): ordered behind the work queued on `stream`, needs no SM. */
int flash_attn_stream_write_flag(int* flag, int value, void* stream);
/* The reference's flash_attention_splitk_merge (FA.cu:559-598; defined there, never launched): merges
 * `splits` independent partial states, o_partial [splits][rows][D] and ml [splits][rows][2] (each written
 * by a flash_attn_fwd_ex call with accumulate = 0), into FP16 O = sum_s w_s O_s / sum_s w_s l_s,
 * w_s = exp(m_s - max m).  One write-only partial per K/V block plus one merge is cheaper than a
 * read-modify-write of the running state per block. */
int flash_attn_merge(const float* o_partial, const float* ml, void* o, int splits, long long rows, int D,
                     void* stream);

/* The reference harness's call pattern with HOST buffers (FA.cu:771-780): H2D of Q,K,V,
 * dispatch, D2H of O, on a per-device cached staging workspace, pipelined over head chunks.  When `ho` is pinned and
 * mapped (cudaHostAlloc / cudaHostRegister), the kernel's epilogue stores the O tiles straight into it with TMA -- no
 * staging of O, no D2H copy; pageable `ho` is copied back from the staging buffer.  FLASH_ATTN_B200_HOST_ZEROCOPY=0
 * forces the copy.  Blocks until O is on the host. */
int flash_attn_fwd_host(const void* hq, const void* hk, const void* hv, void* ho, int B, int H, int N,
                        int D, int causal);

/* Resource report of the kernel flash_attn_fwd would launch for this shape (the reference prints
 * regs/spill/occupancy for its instantiations, FA.cu:711-755). */
typedef struct {
    int regs_per_thread;
    int local_bytes_per_thread; /* spills */
    int static_smem_bytes;
    int dynamic_smem_bytes;
    int threads_per_cta;
    int ctas; /* persistent grid size */
    int tmem_columns;
    int kv_stages; /* K ring entries + V ring entries */
    int work_items; /* (b, h, q-unit) items the grid loops over: 256 query rows each for the default kernel */
    int num_sms;
    int cta_group; /* CTAs cooperating on one item: 1 (default kernel); 2 = tcgen05.mma.cta_group::2 pairs of the
                      experimental kernel selected with FLASH_ATTN_B200_KERNEL=pair */
} flash_attn_kernel_info;
int flash_attn_get_kernel_info(int B, int H, int N, int D, int causal, flash_attn_kernel_info* info);

/* The kernel is persistent: one CTA per SM, all of the SM's shared memory.  A communication kernel
 * (NCCL send/recv of the next K/V block in ring context parallelism) cannot become resident next to
 * it, so its transfer would serialise behind the attention kernel instead of overlapping it.
 * `sms` > 0 makes every later launch of this process leave that many SMs free.  Returns the
 * previous value; 0 (default) uses the whole device. */
int flash_attn_set_sm_margin(int sms);

/* Peer-readable K/V blocks: context parallelism without a communication kernel.
 * With one process per GPU behind NVSwitch every rank can read every other rank's HBM, so the K/V
 * blocks of a long sequence never have to travel around a ring of send/recv kernels: each rank keeps
 * its block where it is, in memory obtained from flash_attn_peer_alloc, publishes the 64-byte handle,
 * and the others map it (flash_attn_peer_open) and PULL the block they need next with a copy engine
 * (flash_attn_peer_copy = one DMA over NVLink, no SM, no shared memory) while the persistent attention
 * grid keeps all 148 SMs.  `handle` is a cudaIpcMemHandle_t; it is valid in other processes of the
 * same node only.  peer_close unmaps an opened block, peer_free releases an allocated one. */
#define FA_PEER_HANDLE_BYTES 64
int flash_attn_peer_alloc(size_t bytes, void** ptr, unsigned char* handle);
int flash_attn_peer_open(const unsigned char* handle, void** ptr);
int flash_attn_peer_close(void* ptr);
int flash_attn_peer_free(void* ptr);
int flash_attn_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
/* the same for `height` rows of `width` bytes with different pitches on either side (a [heads][rows][D] chunk into / out of a
 * buffer that holds more rows per head) */
int flash_attn_peer_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                            void* stream);

/* Kernel watchdog.  Every barrier wait inside the kernel gives up after 10 s (a protocol bug, or a device so
 * oversubscribed that a CTA did not run for that long): the kernel then drains, its output is garbage, and a record
 * {aborted, barrier tag, block, thread} is left for the host.  The launcher reads it without synchronising at the
 * start of every flash_attn_fwd / _bf16 / _ex call (and at the end of flash_attn_fwd_host): the first call after the
 * abort returns FA_ERR_WATCHDOG instead of launching, clears the record and re-arms the device, so later calls work.
 * flash_attn_status() shows the pending record (out4[0] = 1) or the last reported one (out4[0] = 0) of the current
 * device; it neither synchronises nor clears.  (The reference has no such path: its kernel cannot hang, and a CUDA
 * error ends the process, FA.cu:22-30.) */
int flash_attn_status(unsigned int* out4);

/* Number of kernels this library has launched in the calling process (all threads). */
unsigned long long flash_attn_launch_count(void);

/* Frees the per-device state (scheduler words, watchdog mirror, descriptor cache, staging buffers and streams of
 * flash_attn_fwd_host) after synchronising each device; the next call sets a device up again. */
void flash_attn_destroy(void);

const char* flash_attn_error_string(int code);
const char* flash_attn_version(void);

#ifdef __cplusplus
} /* extern "C" */

/* C++ shim with the reference's exact 12-argument signature and error behaviour (print
 * file:line-style message, exit(EXIT_FAILURE); FA.cu:22-30, 606-611).  The split-K buffers are
 * accepted and ignored, as in the reference (never dereferenced there, FA.cu:634-660). */
#if defined(__CUDACC__) || defined(FLASH_ATTN_WITH_CUDA_TYPES)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
static inline void flash_attention_b200_dispatch(const half* Q, const half* K, const half* V,
                                                 half* Output, float* /*splitk_buf_O*/,
                                                 float* /*splitk_buf_ml*/, int batch_size,
                                                 int num_heads, int seq_len, int head_dim,
                                                 bool causal, cudaStream_t stream = 0) {
    int rc = flash_attn_fwd(Q, K, V, Output, batch_size, num_heads, seq_len, head_dim, causal ? 1 : 0,
                            (void*)stream);
    if (rc != FA_OK) {
        fprintf(stderr, "flash_attn_fwd: %s (%d)\n", flash_attn_error_string(rc), rc);
        exit(EXIT_FAILURE);
    }
}
#endif
#endif

#endif /* FLASH_ATTN_B200_H */
