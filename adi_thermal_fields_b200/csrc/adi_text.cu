// adi_text.cu -- the ASCII output path on the device: a C-order (nx,ny,nz) field is
// turned into the text of the reference's VTK writers (vtk_writer.py:4-30 "%.6e", nine values per
// line; waam_from_stl_v7_mm.py:186-215 "%.6g", one line per (k,j) row) by the GPU, byte for byte,
// and streamed to the file in plane chunks:  measure -> scan -> format kernels, text D2H into pinned
// double buffers, write(2) of chunk c while chunk c+1 is formatted and copied.
//
// Work unit = "piece": up to 256 consecutive x cells of one (j,k) row, i.e. 256 consecutive values
// of the file (x is the fastest file index, z the contiguous memory index).  One warp owns a piece:
// rounds each value (adi_fmt_core.h), warp-scans the text lengths, assembles the piece in shared
// memory at the destination's 16-byte phase and stores it with 16-byte vectors.  A block covers
// 16 z planes of the same (x tile, j) so that every 128-byte line of the field is fetched once.
// Bytes: 8 read + ~13 written per value; the reference spends ~1 us of Python per value.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstring>
#include <mutex>

#include "adi_ctx.h"
#include "adi_fmt_core.h"

namespace {

__device__ const double g_pow10[][2] = {ADI_POW10_TABLE};

constexpr int IT = 256;                      // values per piece
constexpr int KT = 16;                       // z planes per block
constexpr int WARPS = 8;
constexpr int SLOT = adifmt::MAX_TEXT + 1;   // worst-case bytes per value incl. separator
constexpr int STAGE = 16 + IT * SLOT + 16;   // per-warp staging bytes (multiple of 16)
static_assert(STAGE % 16 == 0, "stage alignment");

struct TextArgs {
    const void *src;
    int nx, ny, nz;
    int k0, kc;  // planes [k0, k0+kc) of the field are formatted
    int nit;     // pieces per row
    unsigned long long N;
};

template <typename T, int FMT, bool EMIT>
__global__ void __launch_bounds__(WARPS * 32)
k_text(const TextArgs a, uint32_t *__restrict__ piece_len, const uint32_t *__restrict__ piece_off,
       char *__restrict__ out)
{
    __shared__ __align__(16) char stage[EMIT ? WARPS * STAGE : 16];
    constexpr int P = FMT == adifmt::FMT_E6 ? 7 : 6;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int itile = blockIdx.x, j = blockIdx.y, kt = blockIdx.z;
    const double *tab = &g_pow10[0][0];
    const T *__restrict__ src = (const T *)a.src;
    char *st = stage + (EMIT ? warp * STAGE : 0);

    for (int kk = warp; kk < KT; kk += WARPS) {
        const int kl = kt * KT + kk;
        if (kl >= a.kc) break;
        const int k = a.k0 + kl;
        const size_t pid = ((size_t)kl * a.ny + j) * a.nit + itile;
        uint32_t off = 0;
        int phase = 0;
        if (EMIT) {
            off = piece_off[pid];
            phase = (int)(off & 15u);
        }
        int total = 0;
        for (int s = 0; s < IT / 32; ++s) {
            const int i = itile * IT + s * 32 + lane;
            const bool valid = i < a.nx;
            adifmt::Dec d;
            int len = 0;
            if (valid) {
                const double v = (double)src[((size_t)i * a.ny + j) * a.nz + k];
                d = adifmt::round_sig<P>(v, tab);
                len = adifmt::text_len<FMT>(d) + 1;
            }
            int inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int tot = __shfl_sync(0xffffffffu, inc, 31);
            if (EMIT && valid) {
                char *p = st + phase + total + inc - len;
                p = adifmt::emit<FMT>(d, p);
                *p = adifmt::separator(FMT, ((unsigned long long)k * a.ny + j) * a.nx + i, i, a.nx, a.N);
            }
            total += tot;
            if (itile * IT + (s + 1) * 32 >= a.nx) break;
        }
        if (!EMIT) {
            if (lane == 0) piece_len[pid] = (uint32_t)total;
        } else {
            __syncwarp();
            char *dst = out + (off - (uint32_t)phase);  // 16-byte aligned
            const int end = phase + total;
            for (int u = lane; u * 16 < end; u += 32) {
                const int lo = u * 16;
                if (lo >= phase && lo + 16 <= end) {
                    *reinterpret_cast<uint4 *>(dst + lo) = *reinterpret_cast<const uint4 *>(st + lo);
                } else {
                    const int b1 = min(lo + 16, end);
                    for (int b = max(lo, phase); b < b1; ++b) dst[b] = st[b];
                }
            }
            __syncwarp();
        }
    }
}

// exclusive scan of the piece lengths (one block; a chunk has at most a few 1e5 pieces)
__global__ void __launch_bounds__(1024)
k_text_scan(const uint32_t *__restrict__ len, uint32_t *__restrict__ off, size_t n, unsigned long long *total)
{
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x;
    const size_t per = (n + 1023) / 1024;
    const size_t b0 = (size_t)t * per, b = b0 < n ? b0 : n, e = b + per < n ? b + per : n;
    unsigned long long s = 0;
    for (size_t i = b; i < e; ++i) s += len[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long v = t >= o ? part[t - o] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - s;
    for (size_t i = b; i < e; ++i) {
        off[i] = (uint32_t)run;
        run += len[i];
    }
    if (t == 1023) *total = part[1023];
}

// box [lo, hi) of a C-order (nx,ny,nz) array of elem-byte items -> contiguous C-order buffer
template <typename T>
__global__ void k_gather_box(const T *__restrict__ src, T *__restrict__ dst, int ny, int nz, int x0, int y0,
                             int z0, int bx, int by, int bz)
{
    const size_t n = (size_t)bx * by * bz;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(t % bz);
        const size_t r = t / bz;
        const int y = (int)(r % by), x = (int)(r / by);
        dst[t] = src[((size_t)(x0 + x) * ny + (y0 + y)) * nz + (z0 + z)];
    }
}

}  // namespace

namespace adi {

struct ProbeSlot {
    void *d_buf = nullptr, *h_buf = nullptr;
    size_t bytes = 0;
    cudaEvent_t packed = nullptr, landed = nullptr;
    bool busy = false;
};

struct TextState {
    std::mutex mu;   // the text pipeline (buffers, stream, events)
    std::mutex pmu;  // the probe slots: a probe recorded by the stepping thread must not wait for a frame
                     // that a second thread is writing
    std::atomic<long> launches{0};
    uint32_t *d_len = nullptr, *d_off = nullptr;
    size_t pieces_cap = 0;
    unsigned long long *d_total = nullptr, *h_total = nullptr;  // two entries each
    char *d_text[2] = {nullptr, nullptr}, *h_text[2] = {nullptr, nullptr};
    size_t text_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr, ev_total[2] = {nullptr, nullptr}, ev_text[2] = {nullptr, nullptr};
    // probes
    std::vector<ProbeSlot> slots;
    size_t slot_cap = 0;
    cudaStream_t copy_stream = nullptr;
};

static std::mutex g_text_create_mu;

static int text_state(adi_ctx *ctx, TextState **out)
{
    std::lock_guard<std::mutex> create_lock(g_text_create_mu);
    if (!ctx->text) {
        TextState *ts = new TextState();
        ctx->text = ts;  // released by text_release() even if one of the calls below fails
        ADI_CUDA(cudaSetDevice(ctx->device));
        ADI_CUDA(cudaStreamCreateWithFlags(&ts->stream, cudaStreamNonBlocking));
        ADI_CUDA(cudaStreamCreateWithFlags(&ts->copy_stream, cudaStreamNonBlocking));
        ADI_CUDA(cudaEventCreateWithFlags(&ts->ready, cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b) {
            ADI_CUDA(cudaEventCreateWithFlags(&ts->ev_total[b], cudaEventDisableTiming));
            ADI_CUDA(cudaEventCreateWithFlags(&ts->ev_text[b], cudaEventDisableTiming));
        }
        ADI_CUDA(cudaMalloc(&ts->d_total, 2 * sizeof(unsigned long long)));
        ADI_CUDA(cudaMallocHost(&ts->h_total, 2 * sizeof(unsigned long long)));
    }
    *out = ctx->text;
    return ADI_OK;
}

void text_release(adi_ctx *ctx)
{
    TextState *ts = ctx->text;
    if (!ts) return;
    cudaFree(ts->d_len);
    cudaFree(ts->d_off);
    cudaFree(ts->d_total);
    cudaFreeHost(ts->h_total);
    for (int b = 0; b < 2; ++b) {
        cudaFree(ts->d_text[b]);
        cudaFreeHost(ts->h_text[b]);
        if (ts->ev_total[b]) cudaEventDestroy(ts->ev_total[b]);
        if (ts->ev_text[b]) cudaEventDestroy(ts->ev_text[b]);
    }
    for (ProbeSlot &s : ts->slots) {
        cudaFree(s.d_buf);
        cudaFreeHost(s.h_buf);
        if (s.packed) cudaEventDestroy(s.packed);
        if (s.landed) cudaEventDestroy(s.landed);
    }
    if (ts->ready) cudaEventDestroy(ts->ready);
    if (ts->stream) cudaStreamDestroy(ts->stream);
    if (ts->copy_stream) cudaStreamDestroy(ts->copy_stream);
    delete ts;
    ctx->text = nullptr;
}

long text_launches(adi_ctx *ctx) { return ctx->text ? ctx->text->launches.load() : 0; }

static int ensure_pieces(TextState *ts, size_t pieces)
{
    if (pieces <= ts->pieces_cap) return ADI_OK;
    cudaFree(ts->d_len);
    cudaFree(ts->d_off);
    ts->d_len = ts->d_off = nullptr;
    ts->pieces_cap = 0;
    ADI_CUDA(cudaMalloc(&ts->d_len, pieces * sizeof(uint32_t)));
    ADI_CUDA(cudaMalloc(&ts->d_off, pieces * sizeof(uint32_t)));
    ts->pieces_cap = pieces;
    return ADI_OK;
}

template <typename T, int FMT>
static void launch_pass(bool emit, const TextArgs &a, dim3 grid, TextState *ts, char *d_text, cudaStream_t st)
{
    if (emit) k_text<T, FMT, true><<<grid, WARPS * 32, 0, st>>>(a, nullptr, ts->d_off, d_text);
    else k_text<T, FMT, false><<<grid, WARPS * 32, 0, st>>>(a, ts->d_len, nullptr, nullptr);
}

static int launch_text(bool emit, int dtype, int fmt, const TextArgs &a, TextState *ts, char *d_text, cudaStream_t st)
{
    const dim3 grid((unsigned)a.nit, (unsigned)a.ny, (unsigned)((a.kc + KT - 1) / KT));
#define ADI_TEXT_CASE(T)                                                       \
    do {                                                                       \
        if (fmt == adifmt::FMT_E6) launch_pass<T, adifmt::FMT_E6>(emit, a, grid, ts, d_text, st); \
        else launch_pass<T, adifmt::FMT_G6>(emit, a, grid, ts, d_text, st);    \
    } while (0)
    if (dtype == 0) ADI_TEXT_CASE(double);
    else if (dtype == 1) ADI_TEXT_CASE(float);
    else ADI_TEXT_CASE(uint8_t);
#undef ADI_TEXT_CASE
    ts->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

// measure + scan (+ format when `emit`) of planes [k0, k0+kc) on `st`; total -> d_total[slot]
static int enqueue_chunk(TextState *ts, const void *d_field, int dtype, int nx, int ny, int nz, int k0, int kc,
                         int fmt, char *d_text, int slot, bool emit, cudaStream_t st)
{
    TextArgs a;
    a.src = d_field;
    a.nx = nx; a.ny = ny; a.nz = nz;
    a.k0 = k0; a.kc = kc;
    a.nit = (nx + IT - 1) / IT;
    a.N = (unsigned long long)nx * ny * nz;
    const size_t pieces = (size_t)kc * ny * a.nit;
    int rc = launch_text(false, dtype, fmt, a, ts, nullptr, st);
    if (rc) return rc;
    k_text_scan<<<1, 1024, 0, st>>>(ts->d_len, ts->d_off, pieces, ts->d_total + slot);
    ts->launches++;
    ADI_CUDA(cudaGetLastError());
    if (emit) return launch_text(true, dtype, fmt, a, ts, d_text, st);
    return ADI_OK;
}

static int check_field_args(const char *who, const void *d_field, int dtype, int nx, int ny, int nz, int fmt)
{
    if (!d_field || nx <= 0 || ny <= 0 || nz <= 0 || ny > 65535 || dtype < 0 || dtype > 2 ||
        (fmt != adifmt::FMT_E6 && fmt != adifmt::FMT_G6)) {
        set_error(std::string(who) + ": bad field / dtype / format argument");
        return ADI_EINVAL;
    }
    return ADI_OK;
}

// planes per chunk: worst-case text of a chunk stays under 128 MiB (32-bit piece offsets, modest pinned buffers)
static int chunk_planes(int nx, int ny, int nz)
{
    const size_t plane = (size_t)nx * ny;
    size_t kc = ((size_t)128 << 20) / SLOT / plane;
    if (kc >= (size_t)KT) kc -= kc % KT;
    return (int)std::max<size_t>(1, std::min<size_t>(kc, (size_t)nz));
}

}  // namespace adi

extern "C" {

size_t adi_text_capacity(size_t nvalues) { return nvalues * SLOT + 32; }

int adi_text_format(adi_ctx *ctx, const void *d_field, int dtype, int nx, int ny, int nz, int k0, int kc,
                    int fmt, char *d_text, size_t capacity, unsigned long long *nbytes, void *stream)
{
    if (!ctx || !d_text || !nbytes) return ADI_EINVAL;
    int rc = adi::check_field_args("adi_text_format", d_field, dtype, nx, ny, nz, fmt);
    if (rc) return rc;
    if (k0 < 0 || kc <= 0 || k0 + kc > nz || ((uintptr_t)d_text & 15u)) {
        adi::set_error("adi_text_format: plane range outside the field, or d_text not 16-byte aligned");
        return ADI_EINVAL;
    }
    if ((size_t)nx * ny * kc * SLOT >= ((size_t)1 << 32)) {
        adi::set_error("adi_text_format: chunk too large for 32-bit text offsets (format fewer planes per call)");
        return ADI_EINVAL;
    }
    adi::TextState *ts;
    if ((rc = adi::text_state(ctx, &ts))) return rc;
    std::lock_guard<std::mutex> lock(ts->mu);
    ADI_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = adi::ensure_pieces(ts, (size_t)kc * ny * ((nx + IT - 1) / IT)))) return rc;
    if ((rc = adi::enqueue_chunk(ts, d_field, dtype, nx, ny, nz, k0, kc, fmt, nullptr, 0, false, st))) return rc;
    ADI_CUDA(cudaMemcpyAsync(ts->h_total, ts->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ADI_CUDA(cudaStreamSynchronize(st));
    *nbytes = ts->h_total[0];
    if (*nbytes > capacity) {
        adi::set_error("adi_text_format: text buffer too small");
        return ADI_ENOMEM;
    }
    TextArgs a;
    a.src = d_field;
    a.nx = nx; a.ny = ny; a.nz = nz;
    a.k0 = k0; a.kc = kc;
    a.nit = (nx + IT - 1) / IT;
    a.N = (unsigned long long)nx * ny * nz;
    if ((rc = adi::launch_text(true, dtype, fmt, a, ts, d_text, st))) return rc;
    ADI_CUDA(cudaStreamSynchronize(st));
    return ADI_OK;
}

int adi_text_write(adi_ctx *ctx, const char *path, int append, const void *prefix, size_t prefix_len,
                   const void *d_field, int dtype, int nx, int ny, int nz, int fmt,
                   unsigned long long *bytes_written, void *stream)
{
    if (!ctx || !path) return ADI_EINVAL;
    int rc = adi::check_field_args("adi_text_write", d_field, dtype, nx, ny, nz, fmt);
    if (rc) return rc;
    adi::TextState *ts;
    if ((rc = adi::text_state(ctx, &ts))) return rc;
    std::lock_guard<std::mutex> lock(ts->mu);
    ADI_CUDA(cudaSetDevice(ctx->device));

    const int kc = adi::chunk_planes(nx, ny, nz);
    const int nchunks = (nz + kc - 1) / kc;
    const size_t cap = adi_text_capacity((size_t)nx * ny * kc);
    if ((rc = adi::ensure_pieces(ts, (size_t)kc * ny * ((nx + IT - 1) / IT)))) return rc;
    if (cap > ts->text_cap) {
        for (int b = 0; b < 2; ++b) {
            cudaFree(ts->d_text[b]);
            cudaFreeHost(ts->h_text[b]);
            ts->d_text[b] = ts->h_text[b] = nullptr;
        }
        ts->text_cap = 0;
        for (int b = 0; b < 2; ++b) {
            ADI_CUDA(cudaMalloc(&ts->d_text[b], cap));
            ADI_CUDA(cudaMallocHost(&ts->h_text[b], cap));
        }
        ts->text_cap = cap;
    }

    // one write(2) stream: buffered writes to one file serialise on the inode lock (8 pwrite threads
    // measured no faster), so the page cache's ~3.5 GB/s is the ceiling of this path; formatting
    // (~4 ms for 512^3) and the PCIe copy (~35 ms) hide behind it
    const int fd = open(path, O_WRONLY | O_CREAT | (append ? O_APPEND : O_TRUNC), 0644);
    if (fd < 0) {
        adi::set_error(std::string("adi_text_write: cannot open ") + path + ": " + strerror(errno));
        return ADI_EINVAL;
    }
    unsigned long long written = 0;
    auto put = [&](const char *p, size_t n) -> bool {
        while (n) {
            const ssize_t w = write(fd, p, n);
            if (w < 0) {
                if (errno == EINTR) continue;
                return false;
            }
            p += w;
            n -= (size_t)w;
            written += (unsigned long long)w;
        }
        return true;
    };
    auto fail = [&](int code, const std::string &msg) {
        cudaStreamSynchronize(ts->stream);
        close(fd);
        if (!msg.empty()) adi::set_error(msg);
        return code;
    };
#define ADI_TEXT_CUDA(call)                                                                 \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) { adi::cuda_fail(_e, #call); return fail(ADI_ECUDA, ""); }   \
    } while (0)

    if (prefix_len && !put((const char *)prefix, prefix_len))
        return fail(ADI_EINVAL, std::string("adi_text_write: write failed: ") + strerror(errno));

    // the field must be complete on the caller's stream before the writer's stream reads it
    cudaStream_t st = ts->stream;
    ADI_TEXT_CUDA(cudaEventRecord(ts->ready, (cudaStream_t)stream));
    ADI_TEXT_CUDA(cudaStreamWaitEvent(st, ts->ready, 0));

    auto kernels = [&](int c) -> int {  // chunk c: measure, scan, format into d_text[c&1]; total -> h_total[c&1]
        const int b = c & 1, k0 = c * kc, kn = std::min(kc, nz - k0);
        int r = adi::enqueue_chunk(ts, d_field, dtype, nx, ny, nz, k0, kn, fmt, ts->d_text[b], b, true, st);
        if (r) return r;
        if (cudaMemcpyAsync(ts->h_total + b, ts->d_total + b, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            st) != cudaSuccess || cudaEventRecord(ts->ev_total[b], st) != cudaSuccess)
            return ADI_ECUDA;
        return ADI_OK;
    };
    auto copy_out = [&](int c) -> int {  // after the chunk's size is known: text D2H into h_text[c&1]
        const int b = c & 1;
        if (cudaEventSynchronize(ts->ev_total[b]) != cudaSuccess) return ADI_ECUDA;
        if (ts->h_total[b] > cap) return ADI_ENOMEM;
        if (cudaMemcpyAsync(ts->h_text[b], ts->d_text[b], ts->h_total[b], cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(ts->ev_text[b], st) != cudaSuccess)
            return ADI_ECUDA;
        return ADI_OK;
    };

    if ((rc = kernels(0)) || (rc = copy_out(0))) return fail(rc, rc == ADI_ECUDA ? "adi_text_write: CUDA failure" : "");
    if (nchunks > 1 && (rc = kernels(1))) return fail(rc, "adi_text_write: CUDA failure");
    for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        // the other pinned buffer is free (chunk c-1 was written): start moving chunk c+1 now, so that
        // its copy and the kernels of chunk c+2 overlap the write of chunk c
        if (c + 1 < nchunks && (rc = copy_out(c + 1))) return fail(rc, "adi_text_write: CUDA failure");
        ADI_TEXT_CUDA(cudaEventSynchronize(ts->ev_text[b]));
        const unsigned long long n = ts->h_total[b];
        if (c + 2 < nchunks && (rc = kernels(c + 2))) return fail(rc, "adi_text_write: CUDA failure");
        if (!put(ts->h_text[b], (size_t)n))
            return fail(ADI_EINVAL, std::string("adi_text_write: write failed: ") + strerror(errno));
    }
#undef ADI_TEXT_CUDA
    ADI_CUDA(cudaStreamSynchronize(st));
    if (close(fd) != 0) {
        adi::set_error(std::string("adi_text_write: close failed: ") + strerror(errno));
        return ADI_EINVAL;
    }
    if (bytes_written) *bytes_written = written;
    return ADI_OK;
}

/* ---- probes: asynchronous download of lines / slices / boxes ------------------------------- */

int adi_probe_open(adi_ctx *ctx, int nslots, size_t slot_bytes)
{
    if (!ctx || nslots <= 0 || slot_bytes == 0) return ADI_EINVAL;
    adi::TextState *ts;
    int rc = adi::text_state(ctx, &ts);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(ts->pmu);
    ADI_CUDA(cudaSetDevice(ctx->device));
    for (adi::ProbeSlot &s : ts->slots) {
        if (s.busy) {
            adi::set_error("adi_probe_open: a recorded probe has not been fetched yet");
            return ADI_ESTATE;
        }
        cudaFree(s.d_buf);
        cudaFreeHost(s.h_buf);
        if (s.packed) cudaEventDestroy(s.packed);
        if (s.landed) cudaEventDestroy(s.landed);
    }
    ts->slots.assign((size_t)nslots, adi::ProbeSlot());
    ts->slot_cap = slot_bytes;
    for (adi::ProbeSlot &s : ts->slots) {
        ADI_CUDA(cudaMalloc(&s.d_buf, slot_bytes));
        ADI_CUDA(cudaMallocHost(&s.h_buf, slot_bytes));
        ADI_CUDA(cudaEventCreateWithFlags(&s.packed, cudaEventDisableTiming));
        ADI_CUDA(cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming));
    }
    return ADI_OK;
}

int adi_probe_record(adi_ctx *ctx, int slot, const void *d_field, int elem_bytes, int nx, int ny, int nz,
                     const int lo[3], const int hi[3], void *stream)
{
    if (!ctx || !ctx->text || !d_field || !lo || !hi) return ADI_EINVAL;
    adi::TextState *ts = ctx->text;
    std::lock_guard<std::mutex> lock(ts->pmu);
    if (slot < 0 || (size_t)slot >= ts->slots.size()) {
        adi::set_error("adi_probe_record: no such slot (adi_probe_open first)");
        return ADI_EINVAL;
    }
    const int dims[3] = {nx, ny, nz};
    for (int a = 0; a < 3; ++a)
        if (lo[a] < 0 || hi[a] <= lo[a] || hi[a] > dims[a]) {
            adi::set_error("adi_probe_record: empty box or box outside the field");
            return ADI_EINVAL;
        }
    if (elem_bytes != 8 && elem_bytes != 4 && elem_bytes != 1) {
        adi::set_error("adi_probe_record: elem_bytes must be 8, 4 or 1");
        return ADI_EINVAL;
    }
    adi::ProbeSlot &s = ts->slots[(size_t)slot];
    const int bx = hi[0] - lo[0], by = hi[1] - lo[1], bz = hi[2] - lo[2];
    const size_t n = (size_t)bx * by * bz, bytes = n * (size_t)elem_bytes;
    if (bytes > ts->slot_cap) {
        adi::set_error("adi_probe_record: box larger than the slot");
        return ADI_ENOMEM;
    }
    if (s.busy) {
        adi::set_error("adi_probe_record: slot still holds an unfetched record");
        return ADI_ESTATE;
    }
    ADI_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    if (elem_bytes == 8)
        k_gather_box<double><<<grid, 256, 0, st>>>((const double *)d_field, (double *)s.d_buf, ny, nz, lo[0], lo[1],
                                                   lo[2], bx, by, bz);
    else if (elem_bytes == 4)
        k_gather_box<float><<<grid, 256, 0, st>>>((const float *)d_field, (float *)s.d_buf, ny, nz, lo[0], lo[1],
                                                  lo[2], bx, by, bz);
    else
        k_gather_box<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)d_field, (uint8_t *)s.d_buf, ny, nz, lo[0],
                                                    lo[1], lo[2], bx, by, bz);
    ts->launches++;
    ADI_CUDA(cudaGetLastError());
    // the compute stream only pays for the pack kernel; the PCIe copy runs on the copy stream
    ADI_CUDA(cudaEventRecord(s.packed, st));
    ADI_CUDA(cudaStreamWaitEvent(ts->copy_stream, s.packed, 0));
    ADI_CUDA(cudaMemcpyAsync(s.h_buf, s.d_buf, bytes, cudaMemcpyDeviceToHost, ts->copy_stream));
    ADI_CUDA(cudaEventRecord(s.landed, ts->copy_stream));
    s.bytes = bytes;
    s.busy = true;
    return ADI_OK;
}

int adi_probe_fetch(adi_ctx *ctx, int slot, void *h_dst, size_t capacity, size_t *nbytes, int wait)
{
    if (!ctx || !ctx->text || !h_dst) return ADI_EINVAL;
    adi::TextState *ts = ctx->text;
    std::lock_guard<std::mutex> lock(ts->pmu);
    if (slot < 0 || (size_t)slot >= ts->slots.size() || !ts->slots[(size_t)slot].busy) {
        adi::set_error("adi_probe_fetch: nothing recorded in this slot");
        return ADI_ESTATE;
    }
    adi::ProbeSlot &s = ts->slots[(size_t)slot];
    if (s.bytes > capacity) {
        adi::set_error("adi_probe_fetch: destination too small");
        return ADI_ENOMEM;
    }
    if (!wait) {
        const cudaError_t q = cudaEventQuery(s.landed);
        if (q == cudaErrorNotReady) {
            if (nbytes) *nbytes = 0;
            return 1;  /* not there yet */
        }
        ADI_CUDA(q);
    } else {
        ADI_CUDA(cudaEventSynchronize(s.landed));
    }
    memcpy(h_dst, s.h_buf, s.bytes);
    if (nbytes) *nbytes = s.bytes;
    s.busy = false;
    return ADI_OK;
}

}  // extern "C"
