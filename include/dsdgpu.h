/*
 * dsdgpu.h -- C ABI of the B200 (sm_100a) DSD->PCM decimation path.
 *
 * This is the whole boundary between a C++ host (dsf2flac's DsdDecimator, see
 * dsf2flac_b200/host/dsd_decimator.cpp and INTEGRATION.md) and the CUDA side
 * (dsf2flac_b200/csrc/dsdgpu.cu).  Plain pointers and sizes only; no C++ or torch types.
 *
 * The operator it implements is the body of the reference's
 *   DsdDecimator::getSamplesInternal            (reference src/dsd_decimator.cpp:173-222)
 * for frames i = 0..nframes-1 and channels c = 0..nch-1:
 *
 *   p_i  = p0 + i*nstep                                   (index of the newest byte of frame i)
 *   sum  = 0.0; for t = 0..nlut-1: sum += lut[t][ B_c(p_i - t) ]   (fp64, this order)
 *   v    = sum*scale;  if dither_amp>0: v += (r1-r2)*dither_amp;  if clip>0: clamp to +-clip
 *   out[i*nch + c] = (int32) round(v)        (half away from zero)   |  = v   (f64 output)
 *
 * B_c(j) is byte j of channel c of the packed DSD stream: in[c*in_stride + (j - in_first)] (planar
 * rows, the default; see "source_block_bytes" below for the containers' own layouts) for
 * in_first <= j < in_first+in_count and `fill` everywhere else (the reference reader's idle
 * byte 0x69 before the start and after the end of data, dsd_sample_reader.h:122,
 * dsf_file_reader.cpp:94-97).  r1, r2 are the glibc rand() draws number 2m and 2m+1
 * (divided by RAND_MAX), m = dither_first_sample + i*nch + c, i.e. exactly what an unseeded
 * reference process draws for its m-th output sample (dsd_decimator.cpp:201-206).
 *
 * Every function returns 0 on success and a non-zero code on failure; the text is available
 * from dsdgpu_last_error().  There is no CPU fallback: without a usable CUDA device create
 * fails.  A context is not thread-safe; use one per host thread / per GPU.  No global state.
 */
#ifndef DSDGPU_H
#define DSDGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsdgpu_ctx dsdgpu_ctx;

enum {
    DSDGPU_OK = 0,
    DSDGPU_ERR_ARG = 1,      /* bad argument (also: bufferLen not a multiple of nch upstream) */
    DSDGPU_ERR_CUDA = 2,     /* a CUDA runtime call failed */
    DSDGPU_ERR_NODEVICE = 3, /* no CUDA device / device is not usable */
    DSDGPU_ERR_STATE = 4     /* wait without submit, slot busy, ... */
};

/* Output layouts.  INT32 is what the reference hands libFLAC (FLAC__int32 per sample whatever the bit depth,
 * main.cpp:173-177); FLOAT64 is the value before rounding (getSamples<float64>, dsd_decimator.cpp:169-172).
 * PCM24 / PCM16 (the output side) carry the SAME rounded integers as little-endian two's-complement
 * samples of 3 / 2 bytes, interleaved [nframes][nch][3|2]: 25 % / 50 % fewer bytes over PCIe for the 20/24-bit and
 * 16-bit conversions main.cpp runs (clip = 2^(bits-1)-1, main.cpp:245).  They need clip > 0 and clip <= 2^23-1 /
 * 2^15-1, so that every sample fits; the drop-in class widens them back to FLAC__int32 as it hands blocks out. */
enum { DSDGPU_OUT_INT32 = 0, DSDGPU_OUT_FLOAT64 = 1, DSDGPU_OUT_PCM24 = 2, DSDGPU_OUT_PCM16 = 3 };

/* Independent upload/kernel/download pipelines ("slots") a context owns. */
#define DSDGPU_SLOTS 4

/* Kernel selection, for tests and benchmarking (dsdgpu_set_option "kernel"). */
enum {
    DSDGPU_KERNEL_AUTO = 0,   /* ring kernel when the geometry supports it, else direct */
    DSDGPU_KERNEL_DIRECT = 1, /* one thread per output sample, tables [t][byte] in smem */
                              /* 2 was the ring kernel's predecessor (80-step periods); retired */
    DSDGPU_KERNEL_RING = 3,   /* lane-skewed kernel, 72-step periods, 4 frames per lane (divide-by-32) */
    DSDGPU_KERNEL_SEGMENTED = 4 /* lane-skewed kernel for any geometry, <= 72 tables per pass; filters with
                                   more tables (the /64, /128, /256 extension filters) run as a chain of
                                   passes over a plane of running fp64 sums */
};

/* Largest filter a context accepts: 576 tables = the 4607-tap /256 extension filter. */
#define DSDGPU_MAX_TABLES 640

/*
 * Replaces: DsdDecimator::DsdDecimator + initLookupTable (dsd_decimator.cpp:44-72,117-151).
 * The caller builds the nlut x 256 table on the host exactly as initLookupTable does and hands
 * it over; the context uploads it (re-laid out for shared memory) and owns all device state.
 *   device        CUDA device ordinal
 *   nch           channels (>=1)
 *   nlut, nstep   nLookupTable and nStep of the reference (72/4, 30/2, 12/1); any other geometry with
 *                 nlut <= DSDGPU_MAX_TABLES and nstep <= 64 runs on the segmented kernel (extension
 *                 filters for the ratios the reference rejects: 144/8, 288/16, 576/32)
 *   lut           [nlut][256] doubles, row t = table t
 *   lut_precision 64: tables kept in fp64 (bit-exact path).
 *                 32: 32-bit tables -- every entry rounded to a multiple of 2^-30 (fixed point; a
 *                 smaller exponent only if the table's entries could overflow int32 when summed),
 *                 summed exactly and converted once per sample.  Error <= nlut * 2^-31 before
 *                 scaling (3e-8 of full scale, within +-1 LSB at 24 bit); the sums are order
 *                 independent, so every kernel and every sharding returns the same bits; half the
 *                 shared-memory traffic of the fp64 tables
 */
int dsdgpu_create(dsdgpu_ctx** ctx, int device, int nch, int nlut, int nstep, const double* lut,
                  int lut_precision);

/* Replaces: DsdDecimator::~DsdDecimator (dsd_decimator.cpp:74-83). NULL is allowed. */
void dsdgpu_destroy(dsdgpu_ctx* ctx);

/* Last error text of this context (ctx == NULL: of the last failed dsdgpu_create in this thread).
 * Maps to DsdDecimator::getErrorMsg (dsd_decimator.cpp:113-115). */
const char* dsdgpu_last_error(const dsdgpu_ctx* ctx);

/* Knobs: "kernel" (enum above), "chunk_frames" (frames per pipelined host chunk, default 2^20),
 * "slots" (how many of the DSDGPU_SLOTS pipelines dsdgpu_decimate_host cycles through),
 * "source_block_bytes" (layout of `in` / `d_in` for all decimate / submit calls that follow:
 *   0  planar rows, in[c*in_stride + o]                                   o = j - in_first
 *   B  power of two >= 16: DSF data chunk, blocks of B bytes per channel, channel after channel
 *      (dsf_file_reader.cpp:122-149): in[((o / B)*nch + c)*B + o % B]; in_first % B must be 0
 *   1  DSDIFF "DSD " chunk, byte-interleaved (dsdiff_file_reader.cpp:234): in[o*nch + c]
 *   in_stride is ignored for B != 0; the bytes are consumed where they lie, no host re-packing),
 * "dither_stride" (output samples per frame of the WHOLE
 * stream when this context only holds a channel subset of it: sample m of frame i, local channel
 * c is dither_first_sample + i*dither_stride + c; 0 = nch, the default), "dither_cache" (samples, default 2^27 = 1 GiB
 * of device memory at most, allocated as needed; 0 = off: the reference's dither stream comes from the unseeded
 * process-global rand() and dsf2flac converts one file per process (dsd_decimator.cpp:203-204, main.cpp), so the dither
 * term of output sample m is the same constant for every file; a context keeps the part of that plane it has generated
 * and calls whose samples lie below the limit read it in place, generating only what is not there yet), "trace" (1: every
 * dsdgpu_decimate_host call prints the timeline of its chunks -- upload, kernel, download, from events
 * on the slot streams -- to stderr; for looking at the PCIe pipeline, off by default). Unknown keys are
 * an error. */
int dsdgpu_set_option(dsdgpu_ctx* ctx, const char* key, long long value);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
unsigned long long dsdgpu_kernel_launches(const dsdgpu_ctx* ctx);

/* Pinned host memory for the buffers handed to the *_host / submit calls (pageable memory
 * works too, pinned memory is what lets the copies overlap the kernel). */
void* dsdgpu_host_alloc(size_t bytes);
void dsdgpu_host_free(void* p);
/* Page-lock memory the caller already owns (a file payload, an mmap) for the duration of a run. */
int dsdgpu_host_register(void* p, size_t bytes);
int dsdgpu_host_unregister(void* p);

/*
 * Replaces: the loop body of getSamplesInternal (dsd_decimator.cpp:191-221) for a whole run of
 * frames, with every pointer a DEVICE pointer.  Asynchronous on `cuda_stream` (a cudaStream_t
 * passed as void*; NULL = the context's own stream, which is then synchronised before return).
 *   d_in, in_stride, in_first, in_count, fill   the byte source B_c(j) as defined above
 *   p0, nframes                                 first frame position and number of frames
 *   scale, dither_amp, clip                     as main.cpp:239-245 computes them
 *   dither_first_sample                         m of the first output sample (frame-major)
 *   out_kind, d_out                             DSDGPU_OUT_INT32 -> int32[nframes][nch],
 *                                               DSDGPU_OUT_FLOAT64 -> double[nframes][nch],
 *                                               DSDGPU_OUT_PCM24 / _PCM16 -> uint8[nframes][nch][3] / int16[nframes][nch]
 */
int dsdgpu_decimate_device(dsdgpu_ctx* ctx, const uint8_t* d_in, size_t in_stride,
                           int64_t in_first, int64_t in_count, uint8_t fill, int64_t p0,
                           int64_t nframes, double scale, double dither_amp, double clip,
                           uint64_t dither_first_sample, int out_kind, void* d_out,
                           void* cuda_stream);

/*
 * Same operator with HOST buffers: uploads the bytes, runs the kernel, downloads the PCM,
 * internally cut into chunks that cycle through the context's slots so that uploads, kernels and
 * downloads of neighbouring chunks overlap (PCIe runs in both directions at once).  Synchronous.
 * This is what DsdDecimator::getSamples calls (dsd_decimator.cpp:157-159).
 */
int dsdgpu_decimate_host(dsdgpu_ctx* ctx, const uint8_t* in, size_t in_stride, int64_t in_first,
                         int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes,
                         double scale, double dither_amp, double clip,
                         uint64_t dither_first_sample, int out_kind, void* out);

/*
 * The two halves of the above for callers that want to overlap their own work (reader drain,
 * FLAC encode) with the GPU: submit() enqueues upload + kernel + download of one chunk on slot
 * 0..DSDGPU_SLOTS-1 and returns at once; wait() blocks until that slot's output is in `out`.
 * Host buffers must stay untouched between submit and wait.
 */
int dsdgpu_submit(dsdgpu_ctx* ctx, int slot, const uint8_t* in, size_t in_stride,
                  int64_t in_first, int64_t in_count, uint8_t fill, int64_t p0, int64_t nframes,
                  double scale, double dither_amp, double clip, uint64_t dither_first_sample,
                  int out_kind, void* out);
int dsdgpu_wait(dsdgpu_ctx* ctx, int slot);

/*
 * Fills d_out[0..count) (device doubles) with the dither term (r1 - r2) of output samples
 * first_sample .. first_sample+count-1, r = rand()/RAND_MAX of an unseeded glibc process
 * (dsd_decimator.cpp:203-204).  Exposed for tests of the generator.
 */
int dsdgpu_dither_fill_device(dsdgpu_ctx* ctx, uint64_t first_sample, int64_t count, double* d_out,
                              void* cuda_stream);

/*
 * DoP packing -- the other consumer of the same bytes.  Replaces: DopPacker::pack_buffer (reference
 * src/dop_packer.cpp:100-141) for a run of frames:
 *   pm_i = pm0 + 2*i + 1               (the reader's marker after the frame's first step())
 *   out[i*nch + c] = marker(pm_i) + (B_c(pm_i - 1) << 8) + B_c(pm_i)
 *   marker = 0x00050000 if (8*pm_i - 8) % 32 == 0 (C remainder) else 0xFFFA0000
 * with both bytes bit-reversed when reverse_bits != 0 (reader->msbIsPlayedFirst(), :128-131).  pm0 is
 * the reader's marker before the call: -1 right after rewind(), 0 after main.cpp's creep loop
 * (main.cpp:362-364).  B_c, in_first/in_count/fill and the source layouts are those of the decimation
 * calls; the context only supplies the channel count, streams and slots (its tables are not used).
 * Bound: HBM / PCIe, 2 bytes in and 4 bytes out per output sample.
 */
int dsdgpu_dop_pack_device(dsdgpu_ctx* ctx, const uint8_t* d_in, size_t in_stride, int64_t in_first,
                           int64_t in_count, uint8_t fill, int64_t pm0, int64_t nframes, int reverse_bits,
                           int32_t* d_out, void* cuda_stream);
int dsdgpu_dop_pack_host(dsdgpu_ctx* ctx, const uint8_t* in, size_t in_stride, int64_t in_first,
                         int64_t in_count, uint8_t fill, int64_t pm0, int64_t nframes, int reverse_bits,
                         int32_t* out);

/*
 * Multi-GPU partitioning (no reference counterpart: the reference is one thread, its only batch
 * notion is the serial loop of scripts/batch_dsf2flac.sh:16-19).  Output sample (job, channel,
 * frame) depends only on nlut bytes of its own channel (dsd_decimator.cpp:193-198), so a batch
 * of streams is cut into independent shards: by file, by channel, and by time segment with an
 * (nlut-1)-byte input halo.  The batch is laid out as one line of work -- job-major, then
 * channel, then frame, each frame of job k weighing job_cost[k] -- and rank r owns the r-th of
 * `world` equal-cost pieces, cut on multiples of align_frames.  Adjacent channels of one job
 * that cover the same frame range are merged into one shard.  Pure host arithmetic; every rank
 * computes the same plan; there is no data-path collective.
 * Returns the number of shards of `rank` (written to out[0..cap) as far as they fit), or -1 on
 * bad arguments.
 */
typedef struct {
    int32_t job, ch_first, ch_count, reserved;
    int64_t frame_first, frame_count;
} dsdgpu_shard;

int64_t dsdgpu_plan_shards(const int64_t* job_frames, const int32_t* job_nch, const int32_t* job_cost,
                           int njobs, int rank, int world, int64_t align_frames, dsdgpu_shard* out,
                           int64_t cap);

/*
 * A whole batch from C, on the GPUs of this process (no reference counterpart beyond the serial shell loop of
 * scripts/batch_dsf2flac.sh:16-19).  Every stream is described as what DsdDecimator::getSamples would be asked for it
 * (reference src/dsd_decimator.cpp:173-222, driven by main.cpp:169-191): frames i = 0..nframes-1 at byte position
 * p0 + i*nstep, the parameters of main.cpp:239-245, output [nframes][nch].  dsdgpu_plan_shards cuts the batch into
 * world*ndevices equal-cost pieces; this process runs pieces rank*ndevices .. +ndevices-1, one host thread per device,
 * chunks of neighbouring shards and files pipelined through the slots of their contexts (uploads, kernels and downloads
 * overlap across files).  Contexts are kept across calls.  Host buffers should be page-locked (dsdgpu_host_alloc).
 * A shard that holds only some channels of a job needs planar source rows (source_block_bytes = 0); its PCM is put
 * into the job's interleaved output by the host.  Output is bit-identical for every world / ndevices (the dither stream
 * is positional).  Returns 0, or an error code with the text in err.
 */
typedef struct {
    const uint8_t* in;           /* source bytes, laid out as source_block_bytes says (see dsdgpu_set_option) */
    size_t in_stride;            /* planar rows: bytes between channels */
    int64_t in_first, in_count;  /* stream index of the first byte held, bytes per channel held */
    int32_t source_block_bytes;  /* 0 planar, 1 interleaved, 2^k >= 16 block-planar */
    uint8_t fill;                /* byte outside [in_first, in_first + in_count): the reader's idle byte */
    int32_t nch, nlut, nstep, lut_precision;
    const double* lut;           /* [nlut][256], built as initLookupTable builds it (dsd_decimator.cpp:117-151) */
    int64_t p0, nframes;
    double scale, dither_amp, clip;
    uint64_t dither_first_sample; /* m of (frame 0, channel 0): sample (i, c) draws rand() number 2m', 2m'+1, m' = m + i*nch + c */
    int32_t out_kind;            /* DSDGPU_OUT_* */
    void* out;                   /* [nframes][nch] */
} dsdgpu_job;

int dsdgpu_run_batch(const dsdgpu_job* jobs, int njobs, const int* devices, int ndevices, int rank, int world,
                     int64_t align_frames, char* err, size_t err_cap);

#ifdef __cplusplus
}
#endif
#endif /* DSDGPU_H */
