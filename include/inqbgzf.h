/* inqbgzf.h -- C ABI of the GPU BGZF inflate prototype ("GPU inflate later"), part of
 * libinqcall.so.
 *
 * The reference spends its `call` wall time in htslib's BGZF inflate behind `IndexedReader::fetch` / `records()`
 * (src/call.rs:288,294,338,345); the host of this repository is inflate-bound as well. This entry
 * point inflates a batch of BGZF blocks (SAM spec 4.1: raw DEFLATE payloads of <= 64 KB output each) on the
 * device: one warp per block, hand-written sm_100a kernel (inquistr_b200/csrc/inq_inflate.cuh). It is a measured
 * prototype: `inquistr-b200 call` does not use it yet (record parsing would have to move to the device too).
 * There is no CPU implementation behind it; blocks the kernel declines (status != 0) are the caller's to inflate
 * with zlib, and the caller verifies each block's CRC32.
 */
#ifndef INQBGZF_H
#define INQBGZF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one BGZF block: where its raw deflate payload starts inside `comp`, where its output goes inside `out` */
typedef struct inq_zblock {
    uint64_t in_off;
    uint64_t out_off;
    uint32_t in_len;
    uint32_t out_len;       /* ISIZE of the block */
} inq_zblock;

/* status[] values */
#define INQ_Z_OK 0
#define INQ_Z_BAD_CODE 1        /* invalid Huffman code in the stream */
#define INQ_Z_OVERRUN 2         /* output or input overrun */
#define INQ_Z_TABLES 3          /* more long codes than the second-level tables hold: inflate this block on the host */
#define INQ_Z_BAD_HEADER 4
#define INQ_Z_BAD_SIZE 5        /* stream ended before / after out_len bytes */

/*
 * Inflate n_blocks blocks. comp/out/status are HOST pointers (pinned memory makes the copies asynchronous);
 * comp must be readable up to comp_bytes, blocks must not overlap in `out`.
 *   ms_h2d / ms_kernel / ms_d2h (nullable): CUDA-event times of the three phases.
 * Returns 0, or a negative INQ_ERR_* code of include/inqcall.h (message: inq_bgzf_last_error()).
 * A block-level failure is not an error of the call: it is reported in status[].
 */
int inq_bgzf_inflate(int device, const uint8_t *comp, uint64_t comp_bytes, const inq_zblock *blocks, uint32_t n_blocks,
                     uint8_t *out, uint64_t out_bytes, uint32_t *status, float *ms_h2d, float *ms_kernel, float *ms_d2h);
const char *inq_bgzf_last_error(void);

/*
 * Persistent engine for a stream of batches (what `inquistr-b200 call` uses next to its zlib workers): device
 * buffers, stream and events are created once; every run copies the compressed blocks in, inflates, copies the
 * output back and returns when the output is in `out`. comp / out should be page-locked (inq_host_register or
 * inq_host_alloc) so that the copies run at PCIe speed. One engine per host thread; several engines on one device
 * overlap each other's copies and kernels.
 */
typedef struct inq_bgzf_engine inq_bgzf_engine;
int inq_bgzf_engine_create(int device, uint64_t max_comp_bytes, uint64_t max_out_bytes, uint32_t max_blocks, inq_bgzf_engine **out);
void inq_bgzf_engine_destroy(inq_bgzf_engine *eng);
/* blocks[].in_off / out_off are offsets into comp / out; only the byte range [min out_off, max out_off + out_len) of
 * `out` is written, and only [min in_off, max in_off + in_len) of `comp` is read. */
int inq_bgzf_engine_run(inq_bgzf_engine *eng, const uint8_t *comp, uint64_t comp_bytes, const inq_zblock *blocks, uint32_t n_blocks,
                        uint8_t *out, uint32_t *status, float *ms_kernel);
/* page-lock / unlock caller-allocated host memory (cudaHostRegister) */
int inq_host_register(void *p, size_t bytes);
int inq_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif
