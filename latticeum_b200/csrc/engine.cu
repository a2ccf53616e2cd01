// C ABI of the engine (include/lattice_ajtai.h): handle, device buffers, stream plumbing.  No torch types, no
// CPU fallback: every compute entry point ends in a CUDA kernel launch or fails with LAT_E_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lattice_ajtai.h"
#include "kernels.h"

using lat::u64;

namespace {

thread_local std::string g_last_error = "";

int fail_cuda(cudaError_t e, const char *what, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at engine.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
    g_last_error = buf;
    return LAT_E_CUDA;
}
int fail(int status, const std::string &msg) {
    g_last_error = msg;
    return status;
}

#define CK(call)                                                  \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return fail_cuda(_e, #call, __LINE__); \
    } while (0)

constexpr size_t ELEM_BYTES = LAT_RING_DEGREE * sizeof(uint64_t);  // 192

// ---- bounded device-side waits (lat::SpinGuard) ----------------------------------------------------------------------
// One page-locked status word per device, mapped into the device: a kernel whose wait for an upload ticket or for a
// peer's flag times out stores its code there and carries on; the host side turns a non-zero word into LAT_E_CUDA.
constexpr int MAX_DEVICES = 64;
std::mutex g_status_mu;
unsigned long long *g_status_host[MAX_DEVICES] = {}, *g_status_dev[MAX_DEVICES] = {};
std::atomic<unsigned long long> g_spin_timeout_ns{~0ull};  // ~0: not yet read from the environment

unsigned long long spin_timeout_ns() {
    unsigned long long v = g_spin_timeout_ns.load(std::memory_order_relaxed);
    if (v == ~0ull) {
        const char *e = getenv("LAT_SPIN_TIMEOUT_MS");
        v = (e ? strtoull(e, nullptr, 10) : 5000ull) * 1000000ull;
        g_spin_timeout_ns.store(v, std::memory_order_relaxed);
    }
    return v;
}

// the device must be current
int spin_guard(int device, lat::SpinGuard &g) {
    if (device < 0 || device >= MAX_DEVICES) return fail(LAT_E_INVALID_ARGUMENT, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(g_status_mu);
    if (!g_status_host[device]) {
        unsigned long long *hp = nullptr, *dp = nullptr;
        CK(cudaHostAlloc((void **)&hp, sizeof(unsigned long long), cudaHostAllocMapped));
        *hp = 0;
        cudaError_t e = cudaHostGetDevicePointer((void **)&dp, hp, 0);
        if (e != cudaSuccess) {
            cudaFreeHost(hp);
            return fail_cuda(e, "cudaHostGetDevicePointer", __LINE__);
        }
        g_status_host[device] = hp;
        g_status_dev[device] = dp;
    }
    g.status = g_status_dev[device];
    g.timeout_ns = spin_timeout_ns();
    return LAT_OK;
}

// LAT_OK, or LAT_E_CUDA with a message naming the wait that gave up (and the word is cleared)
int spin_status(int device, uint64_t *code_out = nullptr) {
    unsigned long long code = 0;
    if (device >= 0 && device < MAX_DEVICES) {
        std::lock_guard<std::mutex> lock(g_status_mu);
        volatile unsigned long long *hp = g_status_host[device];
        if (hp && (code = *hp)) *hp = 0;
    }
    if (code_out) *code_out = code;
    if (!code) return LAT_OK;
    char buf[256];
    const unsigned long long what = code & 0xff, detail = code >> 8;
    if (what == lat::SPIN_UPLOAD_TICKET)
        snprintf(buf, sizeof(buf), "device %d: a witness kernel gave up waiting for the upload ticket %llu (upload never landed)",
                 device, detail);
    else if (what == lat::SPIN_PEER_FLAG)
        snprintf(buf, sizeof(buf), "device %d: the commitment exchange gave up waiting for peer rank %llu at epoch %llu",
                 device, detail & 0xff, detail >> 8);
    else
        snprintf(buf, sizeof(buf), "device %d: a device-side wait timed out (code %llu)", device, code);
    return fail(LAT_E_CUDA, buf);
}

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t want) {
        if (want <= bytes) return LAT_OK;
        if (p) {
            cudaError_t e = cudaFree(p);
            p = nullptr;
            bytes = 0;
            if (e != cudaSuccess) return fail_cuda(e, "cudaFree", __LINE__);
        }
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail_cuda(e, "cudaMalloc", __LINE__);
        }
        bytes = want;
        return LAT_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};


// ---- per-device scratch for the handle-less host-buffer calls (lat_ring_*, lat_ntt_negacyclic, lat_commitment_sum) ----
// Grow-only device buffers and one non-blocking stream per device, reused across calls: no cudaMalloc / cudaFree and no
// legacy-stream synchronisation on the call path once the buffers have reached their working size.
struct Scratch {
    std::mutex mu;
    DevBuf buf[4];
    cudaStream_t stream = nullptr;
    int *h_flag = nullptr;
};
Scratch g_scratch[MAX_DEVICES];

// locks the device's scratch for the scope, makes the device current, sizes the buffers
struct ScratchLock {
    Scratch *s = nullptr;
    std::unique_lock<std::mutex> lock;
    int open(int device, size_t b0, size_t b1 = 0, size_t b2 = 0, size_t b3 = 0) {
        if (device < 0 || device >= MAX_DEVICES) return fail(LAT_E_INVALID_ARGUMENT, "device ordinal out of range");
        CK(cudaSetDevice(device));
        s = &g_scratch[device];
        lock = std::unique_lock<std::mutex>(s->mu);
        if (!s->stream) CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        if (!s->h_flag) CK(cudaHostAlloc((void **)&s->h_flag, sizeof(int), cudaHostAllocDefault));
        const size_t want[4] = {b0, b1, b2, b3};
        for (int i = 0; i < 4; ++i)
            if (want[i]) {
                int st = s->buf[i].ensure(want[i]);
                if (st) return st;
            }
        return LAT_OK;
    }
    template <class T>
    T *at(int i) const { return s->buf[i].as<T>(); }
    cudaStream_t stream() const { return s->stream; }
};
}  // namespace

struct lat_ajtai {
    int device = 0;
    int sm_count = 148;
    bool mont = false;
    uint32_t kappa = 0, log2_B = 0, L = 0, K = 0;
    u64 n = 0;
    lat::MatLayout lay{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // host-call uploads, pipelined against the kernels
    cudaEvent_t copy_done[4] = {nullptr, nullptr, nullptr, nullptr}, work_done = nullptr;
    cudaEvent_t witness_done = nullptr;          // digits / witness outputs complete: their downloads overlap the matrix-vector kernels
    bool copy_stream_busy = false;               // downloads in flight on copy_stream: finish() waits for them

    DevBuf A;            // re-laid-out matrix, canonical form, 3 words per Fq3 entry (the single-witness kernel)
    // the same matrix as Toom-3 evaluations, 5 words per entry: what the launches with several witnesses read (the K-1
    // planes of a decomposition, commit batches).  Derived on the device from A when such a launch first needs it and
    // again after rows were uploaded; 5/3 of A's bytes, never touched by the single-witness path.
    DevBuf A5;
    lat::MatLayout lay5{};
    uint64_t a_version = 1, a5_version = 0;
    std::vector<uint8_t> row_done;
    uint32_t rows_done = 0;

    DevBuf stage;        // upload staging (one row at a time)
    DevBuf in;           // host-call input staging (w_ccs / f / f_coeff)
    DevBuf f16;          // resident decomposed witness, int16 digits, n x 24
    DevBuf f;            // CRT-form witness, n x 24 u64 (only when a caller wants it on the host)
    DevBuf fx;           // CRT-form witness(es) in the MAC kernel's extended layout, count x n x 48 u64
    DevBuf fx_alt;       // second single-witness buffer for overlapped steps (lat_ajtai_set_step_overlap)
    bool step_overlap = false;
    const void *last_mac_src = nullptr;  // witness buffer the most recent matrix-vector launch reads
    bool last_op_was_mac = false;        // ... and whether that launch is still the newest kernel this handle enqueued
    bool mac_was_last = false;           // snapshot of the above at the start of the current entry point
    DevBuf fcoeff64;     // f_coeff as u64 for host output
    DevBuf planes;       // K x n x 24 CRT-form planes (only when a caller wants them)
    // CRT-form planes of both fold sides, extended layout (MAC input, Toom-3 form), 2K x n x 48 in ONE buffer laid out
    // [side 0: planes 1..K-1][side 1: planes 1..K-1][side 0: plane 0][side 1: plane 0]: the 2 (K-1) planes that get
    // committed are contiguous, so a fold step commits them in one launch (lat_ajtai_fold_step_begin)
    DevBuf planes_all;
    size_t plane_words() const { return (size_t)n * lat::FX_WORDS; }
    int ensure_planes() { return planes_all.ensure((size_t)2 * K * plane_words() * sizeof(u64)); }
    u64 *planes_k1(int side) { return planes_all.as<u64>() + (size_t)side * (K - 1) * plane_words(); }
    u64 *planes_k0(int side) { return planes_all.as<u64>() + ((size_t)2 * (K - 1) + side) * plane_words(); }
    DevBuf planes_lut;   // 48 KB subset-sum table of the planes' transform (planes_kernel), built at first use
    bool side_ready[2] = {false, false};
    int cur_side = 0;
    // fold step (lat_ajtai_fold_step_*): the running accumulator's coefficients as int16 digits, the 2 x K commitments of
    // the current step's two decompositions, the step commitment and the (folded) accumulator commitment
    DevBuf f16_acc, cms_side[2], cm_step, cm_acc;
    bool acc_ready = false, cm_acc_ready = false, fold_pending = false;
    DevBuf rho;          // 2K x 24
    DevBuf f0;           // n x 24, folded witness (host-call staging)
    DevBuf planes_coeff; // K x n x 24 coefficient-form planes (host output only)
    DevBuf cms;          // up to max(K, batch) x kappa x 24
    DevBuf cm_in;        // kappa x 24
    DevBuf ws;           // mac partials
    DevBuf flag;         // int
    int *h_flag = nullptr;  // pinned
    bool has_resident = false;
    // pipelined host-buffer steps (lat_ajtai_submit_w_ccs / lat_ajtai_wait): per-slot input staging, result and flag
    struct Slot {
        DevBuf in, cm, flag, ready;   // w_ccs staging, kappa x 24 result, overflow flag, "upload landed" ticket
        DevBuf out;                   // sharded steps: the full commitment after the exchange
        // page-locked and mapped into the device address space: the last CTA of the matrix-vector kernel writes the
        // commitment, the overflow flag and finally the ticket here; lat_ajtai_wait polls the ticket
        u64 *h_cm = nullptr;
        int *h_flag = nullptr;
        volatile unsigned long long *h_done = nullptr;
        unsigned long long *h_ticket = nullptr;  // source of the ticket copy that follows the upload
        uint64_t ticket = 0;
        uint64_t *user_cm = nullptr;
        bool busy = false;
    };
    Slot slots[LAT_PIPELINE_DEPTH];
    // the same report path for ONE blocking host call: commitment, flag and a sequence number written into mapped host
    // memory by the last CTA, polled by the caller -- no device-to-host copy, no stream synchronisation on the way back
    u64 *blk_cm = nullptr;
    int *blk_flag = nullptr;
    volatile unsigned long long *blk_done = nullptr;
    unsigned long long blk_seq = 0;
    // column sharding of the pipelined steps (lat_ajtai_set_peers)
    bool has_peers = false;
    int peer_rank = 0, peer_world = 1;
    lat::PeerPtrs peers{};
    uint64_t peer_epoch = 1;
    uint64_t next_ticket = 0;
    lat::SpinGuard guard{};  // bound on every device-side wait this handle launches (status word of its device)
    // profiling: pool of event pairs around mac_kernel launches, drained lazily
    static constexpr int EV_POOL = 256;
    bool profiling = false;
    std::vector<cudaEvent_t> ev0, ev1;
    std::vector<uint8_t> ev_pending;
    int ev_next = 0;
    double prof_sum_ms = 0.0;
    uint64_t prof_count = 0;

    int drain_slot(int i) {
        if (!ev_pending[i]) return LAT_OK;
        float ms = 0.f;
        CK(cudaEventSynchronize(ev1[i]));
        CK(cudaEventElapsedTime(&ms, ev0[i], ev1[i]));
        prof_sum_ms += ms;
        prof_count++;
        ev_pending[i] = 0;
        return LAT_OK;
    }

    int bind() {
        CK(cudaSetDevice(device));
        mac_was_last = last_op_was_mac;
        last_op_was_mac = false;
        return LAT_OK;
    }
    int matrix_ready() const {
        if (rows_done != kappa)
            return fail(LAT_E_MATRIX_INCOMPLETE, "Ajtai matrix incomplete: " + std::to_string(rows_done) + " of " +
                                                     std::to_string(kappa) + " rows uploaded");
        return LAT_OK;
    }
    int wrong_len(u64 got) const {
        return fail(LAT_E_WRONG_WITNESS_LENGTH, "Wrong length of the witness: " + std::to_string(got) +
                                                    ", expected: " + std::to_string(n));
    }
    int ensure_a5() {
        if (a5_version == a_version) return LAT_OK;
        const size_t bytes = lay5.total_elems() * sizeof(u64);
        const bool fresh = A5.bytes < bytes;
        int st = A5.ensure(bytes);
        if (st) return st;
        if (fresh) CK(cudaMemsetAsync(A5.p, 0, bytes, stream));  // zero padding columns; rows beyond kappa evaluate to 0
        lat::launch_derive_toom(A.as<u64>(), lay, A5.as<u64>(), lay5, stream);
        CK(cudaGetLastError());
        a5_version = a_version;
        return LAT_OK;
    }
    // mac + reduce into cms_dev for `count` witnesses in the extended layout, count x stride x 48.  toom: Fx is in the
    // Toom-3 form (planes_kernel, fext with toom) and the launch reads the 5-word matrix; else Karatsuba form on A.
    int mac_fx(const u64 *Fx, u64 stride, uint32_t count, u64 *cms_dev, const lat::MacReport &report = lat::MacReport(),
               bool toom = false) {
        int st;
        if (toom && count > 4 && count % 4) {
            // four witnesses per thread is the fastest instance: the largest multiple of 4 first, the rest behind it
            const uint32_t head = count - count % 4;
            if ((st = mac_fx(Fx, stride, head, cms_dev, lat::MacReport(), true))) return st;
            return mac_fx(Fx + (size_t)head * stride * lat::FX_WORDS, stride, count - head,
                          cms_dev + (size_t)head * kappa * LAT_RING_DEGREE, report, true);
        }
        if (toom && (st = ensure_a5())) return st;
        const lat::MatLayout &ml = toom ? lay5 : lay;
        lat::MacPlan plan = lat::plan_mac(ml, count, sm_count);
        size_t had = ws.bytes;
        st = ws.ensure(plan.ws_elems * sizeof(u64));
        if (st) return st;
        if (ws.bytes != had) CK(cudaMemsetAsync(ws.p, 0, ws.bytes, stream));  // the kernel keeps it zero afterwards
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (profiling) {
            int i = ev_next;
            ev_next = (ev_next + 1) % EV_POOL;
            if ((st = drain_slot(i))) return st;  // only blocks if the pool wrapped onto unfinished work
            e0 = ev0[i];
            e1 = ev1[i];
            ev_pending[i] = 1;
        }
        lat::launch_mac(toom ? A5.as<u64>() : A.as<u64>(), ml, Fx, stride, count, plan, ws.as<u64>(), cms_dev, stream, e0, e1, report);
        CK(cudaGetLastError());
        last_mac_src = Fx;
        last_op_was_mac = true;
        return LAT_OK;
    }
    // same for caller-supplied plain CRT-form witnesses (count x stride x 24): extend first
    int mac(const u64 *F, u64 stride, uint32_t count, u64 *cms_dev) {
        int st = fx.ensure((size_t)count * n * lat::FX_WORDS * sizeof(u64));
        if (st) return st;
        const bool toom = count > 1;  // a batch shares every matrix entry between its witnesses
        for (uint32_t p = 0; p < count; ++p)
            lat::launch_fext(F + (size_t)p * stride * LAT_RING_DEGREE, n, fx.as<u64>() + (size_t)p * n * lat::FX_WORDS, stream, toom);
        CK(cudaGetLastError());
        return mac_fx(fx.as<u64>(), n, count, cms_dev, lat::MacReport(), toom);
    }
    int clear_flag() {
        CK(cudaMemsetAsync(flag.p, 0, sizeof(int), stream));
        return LAT_OK;
    }
    // synchronise, fetch and clear the overflow flag
    int finish() {
        CK(cudaMemcpyAsync(h_flag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemsetAsync(flag.p, 0, sizeof(int), stream));
        CK(cudaStreamSynchronize(stream));
        if (copy_stream_busy) {
            copy_stream_busy = false;
            CK(cudaStreamSynchronize(copy_stream));
        }
        int st = spin_status(device);  // a bounded device-side wait gave up: the results are not to be trusted
        if (st) return st;
        if (*h_flag)
            return fail(LAT_E_DIGIT_OVERFLOW, "a coefficient needs more digits than the decomposition padding allows");
        return LAT_OK;
    }
    // pipelined-step state (lat_ajtai_submit_w_ccs): everything a slot needs, allocated and initialised once at
    // creation on the handle's own stream and synchronised there -- nothing is allocated or memset on the submit path
    // (a first-use cudaMemset on the legacy stream once raced the ticket copy on copy_stream and hung the step).
    int slots_init() {
        const size_t cm_bytes = (size_t)kappa * ELEM_BYTES, in_bytes = ((n + L - 1) / L) * ELEM_BYTES;
        int st;
        for (Slot &sl : slots) {
            if ((st = sl.in.ensure(in_bytes)) || (st = sl.cm.ensure(cm_bytes)) || (st = sl.out.ensure(cm_bytes)) ||
                (st = sl.flag.ensure(sizeof(int))) || (st = sl.ready.ensure(sizeof(unsigned long long))))
                return st;
            CK(cudaMemsetAsync(sl.flag.p, 0, sizeof(int), stream));
            CK(cudaMemsetAsync(sl.ready.p, 0xff, sizeof(unsigned long long), stream));  // no ticket has this value
            CK(cudaHostAlloc((void **)&sl.h_cm, cm_bytes, cudaHostAllocMapped));
            CK(cudaHostAlloc((void **)&sl.h_flag, sizeof(int), cudaHostAllocMapped));
            CK(cudaHostAlloc((void **)&sl.h_ticket, sizeof(unsigned long long), cudaHostAllocDefault));
            CK(cudaHostAlloc((void **)&sl.h_done, sizeof(unsigned long long), cudaHostAllocMapped));
            *sl.h_flag = 0;
            *sl.h_done = ~0ull;
        }
        CK(cudaHostAlloc((void **)&blk_cm, cm_bytes, cudaHostAllocMapped));
        CK(cudaHostAlloc((void **)&blk_flag, sizeof(int), cudaHostAllocMapped));
        CK(cudaHostAlloc((void **)&blk_done, sizeof(unsigned long long), cudaHostAllocMapped));
        *blk_flag = 0;
        *blk_done = 0;
        return LAT_OK;
    }
};

extern "C" {

const char *lat_strerror(int status) {
    switch (status) {
        case LAT_OK: return "ok";
        case LAT_E_WRONG_WITNESS_LENGTH: return "wrong length of the witness";
        case LAT_E_WRONG_COMMITMENT_LENGTH: return "wrong length of the commitment";
        case LAT_E_WRONG_MATRIX_DIMENSIONS: return "Ajtai matrix has wrong dimensions";
        case LAT_E_DIGIT_OVERFLOW: return "coefficient does not fit in the requested number of digits";
        case LAT_E_INVALID_ARGUMENT: return "invalid argument";
        case LAT_E_CUDA: return "CUDA failure (no CPU fallback exists)";
        case LAT_E_MATRIX_INCOMPLETE: return "Ajtai matrix rows missing";
        default: return "unknown status";
    }
}
const char *lat_last_error(void) { return g_last_error.c_str(); }
int lat_abi_version(void) { return LAT_ABI_VERSION; }

int lat_ajtai_create(lat_ajtai **out, uint32_t kappa, uint64_t n, uint32_t log2_B, uint32_t L, uint32_t K, int repr,
                     int device) {
    if (!out) return fail(LAT_E_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (kappa == 0 || n == 0) return fail(LAT_E_WRONG_MATRIX_DIMENSIONS, "kappa and n must be positive");
    if (log2_B < 1 || log2_B > 15 || L < 1 || L > 8 || K < 1 || K > 15)
        return fail(LAT_E_INVALID_ARGUMENT, "need 1<=log2_B<=15, 1<=L<=8, 1<=K<=15");
    if (repr != LAT_REPR_CANONICAL && repr != LAT_REPR_MONTGOMERY) return fail(LAT_E_INVALID_ARGUMENT, "bad repr");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LAT_E_CUDA, "no such CUDA device (this engine has no CPU fallback)");
    CK(cudaSetDevice(device));
    lat_ajtai *h = new (std::nothrow) lat_ajtai();
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "out of host memory");
    h->device = device;
    h->mont = repr == LAT_REPR_MONTGOMERY;
    h->kappa = kappa;
    h->n = n;
    h->log2_B = log2_B;
    h->L = L;
    h->K = K;
    h->lay = lat::make_layout(kappa, n);
    h->lay5 = lat::make_layout(kappa, n, true);
    h->row_done.assign(kappa, 0);
    int st = LAT_OK;
    cudaError_t e;
    do {
        if ((e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) break;
        if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        h->stream = h->own_stream;
        if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming);
        if (e != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&h->work_done, cudaEventDisableTiming)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&h->witness_done, cudaEventDisableTiming)) != cudaSuccess) break;
        if ((e = cudaHostAlloc((void **)&h->h_flag, sizeof(int), cudaHostAllocDefault)) != cudaSuccess) break;
        size_t a_bytes = h->lay.total_elems() * sizeof(u64);
        if ((st = h->A.ensure(a_bytes))) break;
        if ((e = cudaMemsetAsync(h->A.p, 0, a_bytes, h->stream)) != cudaSuccess) break;  // zero padding rows/columns
        if ((st = h->flag.ensure(sizeof(int)))) break;
        if ((e = cudaMemsetAsync(h->flag.p, 0, sizeof(int), h->stream)) != cudaSuccess) break;
        if ((st = h->f16.ensure(n * LAT_RING_DEGREE * sizeof(int16_t)))) break;
        if ((st = h->fx.ensure(n * lat::FX_WORDS * sizeof(u64)))) break;
        if ((st = h->cms.ensure((size_t)K * kappa * ELEM_BYTES))) break;
        if ((st = h->cm_in.ensure((size_t)kappa * ELEM_BYTES))) break;
        // first-use allocations would stall (and implicitly synchronise) the kernel chains later: make them here
        if ((st = h->fx_alt.ensure(n * lat::FX_WORDS * sizeof(u64)))) break;
        {
            const uint32_t max_planes = 2 * K > 32 ? 2 * K : 32;
            if ((st = h->ws.ensure(lat::plan_mac(h->lay, max_planes, h->sm_count).ws_elems * sizeof(u64)))) break;
            if ((e = cudaMemsetAsync(h->ws.p, 0, h->ws.bytes, h->stream)) != cudaSuccess) break;
        }
        if ((st = spin_guard(device, h->guard))) break;
        if ((st = h->slots_init())) break;
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) break;
    } while (0);
    if (st == LAT_OK && e != cudaSuccess) st = fail_cuda(e, "lat_ajtai_create", __LINE__);
    if (st != LAT_OK) {
        lat_ajtai_destroy(h);
        return st;
    }
    *out = h;
    return LAT_OK;
}

void lat_ajtai_destroy(lat_ajtai *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    if (h->stream && h->stream != h->own_stream) cudaStreamSynchronize(h->stream);  // steps in flight write into our buffers
    DevBuf *bufs[] = {&h->A, &h->A5, &h->planes_lut, &h->stage, &h->in, &h->f16, &h->f, &h->fx, &h->fcoeff64, &h->planes, &h->planes_all,
                      &h->rho, &h->f0, &h->planes_coeff, &h->cms, &h->cm_in, &h->ws, &h->flag, &h->fx_alt,
                      &h->f16_acc, &h->cms_side[0], &h->cms_side[1], &h->cm_step, &h->cm_acc};
    for (DevBuf *b : bufs) b->release();
    if (h->h_flag) cudaFreeHost(h->h_flag);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (lat_ajtai::Slot &sl : h->slots) {
        sl.in.release();
        sl.cm.release();
        sl.flag.release();
        sl.ready.release();
        sl.out.release();
        if (sl.h_cm) cudaFreeHost(sl.h_cm);
        if (sl.h_flag) cudaFreeHost(sl.h_flag);
        if (sl.h_done) cudaFreeHost((void *)sl.h_done);
        if (sl.h_ticket) cudaFreeHost(sl.h_ticket);
    }
    if (h->blk_cm) cudaFreeHost(h->blk_cm);
    if (h->blk_flag) cudaFreeHost(h->blk_flag);
    if (h->blk_done) cudaFreeHost((void *)h->blk_done);
    for (cudaEvent_t e : h->ev0) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev1) cudaEventDestroy(e);
    for (cudaEvent_t ev : h->copy_done)
        if (ev) cudaEventDestroy(ev);
    if (h->work_done) cudaEventDestroy(h->work_done);
    if (h->witness_done) cudaEventDestroy(h->witness_done);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

uint32_t lat_ajtai_kappa(const lat_ajtai *h) { return h ? h->kappa : 0; }
uint64_t lat_ajtai_width(const lat_ajtai *h) { return h ? h->n : 0; }

int lat_ajtai_set_stream(lat_ajtai *h, void *cuda_stream) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    int st = h->bind();
    if (st) return st;
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return LAT_OK;
}

int lat_ajtai_synchronize(lat_ajtai *h) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    int st = h->bind();
    if (st) return st;
    return h->finish();
}

static int mark_rows(lat_ajtai *h, uint32_t row0, uint32_t nrows) {
    h->a_version++;  // the 5-word copy (ensure_a5) is stale now
    for (uint32_t r = row0; r < row0 + nrows; ++r)
        if (!h->row_done[r]) {
            h->row_done[r] = 1;
            h->rows_done++;
        }
    return LAT_OK;
}

int lat_ajtai_upload_rows_dev(lat_ajtai *h, uint32_t row0, uint32_t nrows, const uint64_t *rows_dev, uint64_t row_stride) {
    if (!h || !rows_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if ((u64)row0 + nrows > h->kappa || row_stride < h->n)
        return fail(LAT_E_WRONG_MATRIX_DIMENSIONS, "rows outside the kappa x n matrix");
    int st = h->bind();
    if (st) return st;
    lat::launch_relayout((const u64 *)rows_dev, row0, nrows, row_stride, h->mont, h->lay, h->A.as<u64>(), h->stream);
    CK(cudaGetLastError());
    return mark_rows(h, row0, nrows);
}

int lat_ajtai_upload_rows(lat_ajtai *h, uint32_t row0, uint32_t nrows, const uint64_t *rows, uint64_t row_stride) {
    if (!h || !rows) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if ((u64)row0 + nrows > h->kappa || row_stride < h->n)
        return fail(LAT_E_WRONG_MATRIX_DIMENSIONS, "rows outside the kappa x n matrix");
    int st = h->bind();
    if (st) return st;
    size_t row_bytes = h->n * ELEM_BYTES;
    if ((st = h->stage.ensure(row_bytes))) return st;
    for (uint32_t r = 0; r < nrows; ++r) {
        // stream order makes the single staging row safe: the copy of row r+1 starts after relayout of row r
        CK(cudaMemcpyAsync(h->stage.p, rows + (size_t)r * row_stride * LAT_RING_DEGREE, row_bytes, cudaMemcpyHostToDevice,
                           h->stream));
        lat::launch_relayout(h->stage.as<u64>(), row0 + r, 1, h->n, h->mont, h->lay, h->A.as<u64>(), h->stream);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    return mark_rows(h, row0, nrows);
}

// ---- commit ----------------------------------------------------------------------------------------------------
int lat_ajtai_commit_ntt_batch_dev(lat_ajtai *h, const uint64_t *fs_dev, uint32_t count, uint64_t f_len,
                                   uint64_t *cms_dev) {
    if (!h || !fs_dev || !cms_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (f_len != h->n) return h->wrong_len(f_len);
    if (count == 0) return LAT_OK;
    int st = h->bind();
    if (st || (st = h->matrix_ready())) return st;
    return h->mac((const u64 *)fs_dev, f_len, count, (u64 *)cms_dev);
}
int lat_ajtai_commit_ntt_dev(lat_ajtai *h, const uint64_t *f_dev, uint64_t f_len, uint64_t *cm_dev) {
    return lat_ajtai_commit_ntt_batch_dev(h, f_dev, 1, f_len, cm_dev);
}

int lat_ajtai_commit_ntt_batch(lat_ajtai *h, const uint64_t *fs, uint32_t count, uint64_t f_len, uint64_t *cms) {
    if (!h || !fs || !cms) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (f_len != h->n) return h->wrong_len(f_len);
    if (count == 0) return LAT_OK;
    int st = h->bind();
    if (st || (st = h->matrix_ready())) return st;
    size_t in_bytes = (size_t)count * f_len * ELEM_BYTES, out_bytes = (size_t)count * h->kappa * ELEM_BYTES;
    if ((st = h->in.ensure(in_bytes)) || (st = h->cms.ensure(out_bytes))) return st;
    CK(cudaMemcpyAsync(h->in.p, fs, in_bytes, cudaMemcpyHostToDevice, h->stream));
    if ((st = h->mac(h->in.as<u64>(), f_len, count, h->cms.as<u64>()))) return st;
    CK(cudaMemcpyAsync(cms, h->cms.p, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    return h->finish();
}
int lat_ajtai_commit_ntt(lat_ajtai *h, const uint64_t *f, uint64_t f_len, uint64_t *cm) {
    return lat_ajtai_commit_ntt_batch(h, f, 1, f_len, cm);
}

int lat_ajtai_commit_coeff(lat_ajtai *h, const uint64_t *f_coeff, uint64_t f_len, uint64_t *cm) {
    if (!h || !f_coeff || !cm) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (f_len != h->n) return h->wrong_len(f_len);
    int st = h->bind();
    if (st || (st = h->matrix_ready())) return st;
    size_t in_bytes = f_len * ELEM_BYTES;
    if ((st = h->in.ensure(in_bytes))) return st;
    CK(cudaMemcpyAsync(h->in.p, f_coeff, in_bytes, cudaMemcpyHostToDevice, h->stream));
    lat::launch_crt(h->in.as<u64>(), h->in.as<u64>(), f_len, h->stream);  // in place
    CK(cudaGetLastError());
    if ((st = h->mac(h->in.as<u64>(), f_len, 1, h->cms.as<u64>()))) return st;
    CK(cudaMemcpyAsync(cm, h->cms.p, (size_t)h->kappa * ELEM_BYTES, cudaMemcpyDeviceToHost, h->stream));
    return h->finish();
}

// shared by from_w_ccs and decompose_and_commit_*: device input -> digits -> CRT -> (commit)
static int witness_core(lat_ajtai *h, const u64 *w_dev, u64 w_len, bool in_coeff, u64 *f_coeff_dev, u64 *f_dev,
                        u64 *cm_dev, const unsigned long long *ready_flag = nullptr, unsigned long long ready_value = 0) {
    // one kernel: iCRT -> digits -> CRT; the extended layout feeds the MAC, the plain layout only if the caller wants f
    // With step overlap this kernel may start while the previous commitment's matrix-vector kernel is draining: it
    // then writes the witness buffer that kernel is NOT reading (the protocol in ring_kernels.cu / mac_kernels.cu
    // makes the buffer of two steps back safe to reuse).  Event brackets (profiling) serialise the launches anyway.
    u64 *fxp = h->fx.as<u64>();
    const bool chained = h->step_overlap && cm_dev && h->mac_was_last && !h->profiling;
    if (chained && h->last_mac_src == h->fx.p) fxp = h->fx_alt.as<u64>();
    h->guard.timeout_ns = spin_timeout_ns();
    lat::launch_witness(w_dev, w_len, (int)h->log2_B, (int)h->L, h->mont, in_coeff, h->f16.as<int16_t>(), f_coeff_dev,
                        f_dev, cm_dev ? fxp : nullptr, h->flag.as<int>(), h->stream, chained, ready_flag, ready_value, h->guard);
    CK(cudaGetLastError());
    h->has_resident = true;
    if (cm_dev) return h->mac_fx(fxp, h->n, 1, cm_dev);
    return LAT_OK;
}

int lat_ajtai_witness_from_w_ccs_dev(lat_ajtai *h, const uint64_t *w_ccs_dev, uint64_t w_len, uint64_t *f_coeff_dev,
                                     uint64_t *f_dev, uint64_t *cm_dev) {
    if (!h || !w_ccs_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (w_len * h->L != h->n) return h->wrong_len(w_len * h->L);
    int st = h->bind();
    if (st) return st;
    if (cm_dev && (st = h->matrix_ready())) return st;
    return witness_core(h, (const u64 *)w_ccs_dev, w_len, false, (u64 *)f_coeff_dev, (u64 *)f_dev, (u64 *)cm_dev);
}

int lat_ajtai_witness_from_w_ccs_gated_dev(lat_ajtai *h, const uint64_t *w_ccs_dev, uint64_t w_len, uint64_t *cm_dev,
                                           const uint64_t *ready_flag_dev, uint64_t ready_value) {
    if (!h || !w_ccs_dev || !cm_dev || !ready_flag_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (w_len * h->L != h->n) return h->wrong_len(w_len * h->L);
    int st = h->bind();
    if (st || (st = h->matrix_ready())) return st;
    return witness_core(h, (const u64 *)w_ccs_dev, w_len, false, nullptr, nullptr, (u64 *)cm_dev,
                        (const unsigned long long *)ready_flag_dev, ready_value);
}

static int device_view(void *host, void **dev) {
    CK(cudaHostGetDevicePointer(dev, host, 0));
    return LAT_OK;
}

// Host-buffer w (w_ccs, or coefficients when in_coeff) -> resident digits (+ optional u64 outputs on the device) and,
// with want_cm, the commitment in cm_dev.  Enqueues only; the caller copies results out and calls finish().
// input_staged: the caller has already copied w into h->in (and made h->stream wait for that copy); w is not touched.
static int witness_enqueue(lat_ajtai *h, const uint64_t *w, uint64_t w_len, bool in_coeff, u64 *d_fc, u64 *d_f, u64 *cm_dev,
                           bool early_downloads = false, const lat::MacReport &report = lat::MacReport(),
                           bool input_staged = false) {
    int st;
    const size_t in_bytes = w_len * ELEM_BYTES;
    if ((st = h->in.ensure(in_bytes))) return st;
    // Pipeline the upload with the witness kernel: w_ccs goes up in chunks on a copy stream and each chunk's
    // iCRT/decompose/CRT starts as soon as its bytes have landed (per-element work, so chunks are independent); only
    // the matrix-vector kernel needs the whole witness.
    u64 *d_fx = cm_dev ? h->fx.as<u64>() : nullptr;
    // Pinned (page-locked) host memory is mapped into the device address space: the witness kernel then reads w_ccs
    // straight over PCIe -- no staging copy, no extra launch, the 8-lane loads are contiguous 768-byte runs per warp.
    const u64 *w_mapped = nullptr;
    if (!input_staged) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, w) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            w_mapped = static_cast<const u64 *>(attr.devicePointer);
        else
            cudaGetLastError();  // pageable memory: not an error, take the copy path
    }
    u64 nchunks = w_len >= 4096 ? 4 : 1;
    if (input_staged) {
        nchunks = 0;
        lat::launch_witness(h->in.as<u64>(), w_len, (int)h->log2_B, (int)h->L, h->mont, in_coeff, h->f16.as<int16_t>(), d_fc, d_f,
                            d_fx, h->flag.as<int>(), h->stream);
        CK(cudaGetLastError());
    } else if (w_mapped) {
        nchunks = 0;
        static const bool no_stage = getenv("LAT_NO_STAGE_INPUT") != nullptr;
        const bool stage = !no_stage && (reinterpret_cast<uintptr_t>(w_mapped) & 15) == 0;  // bulk copies want 16-byte alignment
        lat::launch_witness(w_mapped, w_len, (int)h->log2_B, (int)h->L, h->mont, in_coeff, h->f16.as<int16_t>(), d_fc, d_f,
                            d_fx, h->flag.as<int>(), h->stream, false, nullptr, 0, lat::SpinGuard(), stage);
        CK(cudaGetLastError());
    }
    for (u64 c = 0; c < nchunks; ++c) {
        const u64 e0 = w_len * c / nchunks, e1 = w_len * (c + 1) / nchunks, cnt = e1 - e0, lo = e0 * h->L;
        cudaStream_t cs = nchunks > 1 ? h->copy_stream : h->stream;
        CK(cudaMemcpyAsync(h->in.as<u64>() + e0 * LAT_RING_DEGREE, w + e0 * LAT_RING_DEGREE, cnt * ELEM_BYTES,
                           cudaMemcpyHostToDevice, cs));
        if (nchunks > 1) {
            CK(cudaEventRecord(h->copy_done[c], cs));
            CK(cudaStreamWaitEvent(h->stream, h->copy_done[c], 0));
        }
        lat::launch_witness(h->in.as<u64>() + e0 * LAT_RING_DEGREE, cnt, (int)h->log2_B, (int)h->L, h->mont, in_coeff,
                            h->f16.as<int16_t>() + lo * LAT_RING_DEGREE, d_fc ? d_fc + lo * LAT_RING_DEGREE : nullptr,
                            d_f ? d_f + lo * LAT_RING_DEGREE : nullptr, d_fx ? d_fx + lo * lat::FX_WORDS : nullptr,
                            h->flag.as<int>(), h->stream);
        CK(cudaGetLastError());
    }
    h->has_resident = true;
    if (early_downloads) {
        // the caller downloads the witness outputs (digits, f_coeff, f) on copy_stream while the matrix-vector kernels
        // run: they are complete here.  (The event between the two kernels costs the programmatic overlap of the
        // matrix-vector kernel's prologue, a few microseconds against ~90 us of PCIe transfer hidden.)
        CK(cudaEventRecord(h->witness_done, h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->witness_done, 0));
        h->copy_stream_busy = true;
    }
    if (cm_dev && (st = h->mac_fx(h->fx.as<u64>(), h->n, 1, cm_dev, report))) return st;
    if (nchunks > 1) {  // the next call's copies must not overtake this call's kernels reading h->in
        CK(cudaEventRecord(h->work_done, h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->work_done, 0));
    }
    return LAT_OK;
}

static int witness_host(lat_ajtai *h, const uint64_t *w, uint64_t w_len, bool in_coeff, uint64_t *f_coeff, uint64_t *f,
                        uint64_t *cm, int16_t *f_coeff16 = nullptr) {
    if (!h || !w) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (w_len * h->L != h->n) return h->wrong_len(w_len * h->L);
    int st = h->bind();
    if (st) return st;
    if (cm && (st = h->matrix_ready())) return st;
    const size_t n_bytes = h->n * ELEM_BYTES;
    if (f_coeff && (st = h->fcoeff64.ensure(n_bytes))) return st;
    if (f && (st = h->f.ensure(n_bytes))) return st;
    u64 *d_fc = f_coeff ? h->fcoeff64.as<u64>() : nullptr, *d_f = f ? h->f.as<u64>() : nullptr;
    const bool early = cm && (f_coeff16 || f_coeff || f);
    static const bool no_report = getenv("LAT_NO_BLOCKING_REPORT") != nullptr;
    if (cm && !early && !no_report) {
        // Commitment only: nothing comes back but 6 KB, so the last CTA of the matrix-vector kernel writes it (with the
        // overflow flag and a sequence number) straight into mapped host memory and this thread polls the number -- the
        // way back costs one PCIe write latency instead of a copy launch plus a stream synchronisation.
        lat::MacReport rep;
        unsigned long long *done_dev = nullptr;
        if ((st = device_view(h->blk_cm, (void **)&rep.cm_host)) || (st = device_view(h->blk_flag, (void **)&rep.flag_host)) ||
            (st = device_view(const_cast<unsigned long long *>(h->blk_done), (void **)&done_dev)))
            return st;
        rep.flag_dev = h->flag.as<int>();
        rep.done_host = done_dev;
        const unsigned long long seq = rep.done_value = ++h->blk_seq;
        if ((st = witness_enqueue(h, w, w_len, in_coeff, nullptr, nullptr, h->cms.as<u64>(), false, rep))) return st;
        for (unsigned spins = 0; *h->blk_done != seq; ++spins) {
            if ((spins & 0xffff) == 0xffff) {  // look at the stream now and then: a device fault must not hang the caller
                cudaError_t e = cudaStreamQuery(h->stream);
                if (e != cudaSuccess && e != cudaErrorNotReady) return fail_cuda(e, "lat_ajtai_witness_from_w_ccs", __LINE__);
                if (e == cudaSuccess && *h->blk_done != seq) return fail(LAT_E_CUDA, "stream idle but the call never reported completion");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        if ((st = spin_status(h->device))) return st;
        memcpy(cm, h->blk_cm, (size_t)h->kappa * ELEM_BYTES);
        if (*h->blk_flag)
            return fail(LAT_E_DIGIT_OVERFLOW, "a coefficient needs more digits than the decomposition padding allows");
        return LAT_OK;
    }
    if ((st = witness_enqueue(h, w, w_len, in_coeff, d_fc, d_f, cm ? h->cms.as<u64>() : nullptr, early))) return st;
    if (cm) CK(cudaMemcpyAsync(cm, h->cms.p, (size_t)h->kappa * ELEM_BYTES, cudaMemcpyDeviceToHost, h->stream));
    cudaStream_t ds = early ? h->copy_stream : h->stream;  // with a commitment to compute, the downloads run beside it
    if (f_coeff16) CK(cudaMemcpyAsync(f_coeff16, h->f16.p, h->n * LAT_RING_DEGREE * sizeof(int16_t), cudaMemcpyDeviceToHost, ds));
    if (f_coeff) CK(cudaMemcpyAsync(f_coeff, h->fcoeff64.p, n_bytes, cudaMemcpyDeviceToHost, ds));
    if (f) CK(cudaMemcpyAsync(f, h->f.p, n_bytes, cudaMemcpyDeviceToHost, ds));
    if (early) {  // later calls write the buffers these downloads read: order them behind
        CK(cudaEventRecord(h->work_done, h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->work_done, 0));
    }
    return h->finish();
}

int lat_ajtai_witness_from_w_ccs(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, uint64_t *f_coeff, uint64_t *f,
                                 uint64_t *cm) {
    return witness_host(h, w_ccs, w_len, false, f_coeff, f, cm);
}
int lat_ajtai_witness_from_w_ccs_compact(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, int16_t *f_coeff16,
                                         uint64_t *f, uint64_t *cm) {
    return witness_host(h, w_ccs, w_len, false, nullptr, f, cm, f_coeff16);
}
int lat_ajtai_decompose_and_commit_ntt(lat_ajtai *h, const uint64_t *w, uint64_t w_len, uint64_t *cm) {
    if (!cm) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    return witness_host(h, w, w_len, false, nullptr, nullptr, cm);
}
int lat_ajtai_decompose_and_commit_coeff(lat_ajtai *h, const uint64_t *w_coeff, uint64_t w_len, uint64_t *cm) {
    if (!cm) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    return witness_host(h, w_coeff, w_len, true, nullptr, nullptr, cm);
}

// ---- non-blocking steps -------------------------------------------------------------------------------------------------
// submit() queues upload (copy engine) -> witness kernel -> matrix-vector kernel (-> exchange kernel when sharded) and
// returns; the last kernel writes the commitment and then the ticket into mapped host memory, wait() polls that word.
// The compute stream carries nothing but kernels, so several tickets overlap on the device.  This is for work that is
// independent of the commitment (the host's own work of the same step, independent provers): the IVC steps of one
// zkVM run are strictly dependent -- step i+1's z is built from fold(cm_i, w_i) (ZKVM/main.rs:140-156,174-182) -- so a
// drop-in caller gets one ticket's latency per step, not the overlapped throughput.
int lat_ajtai_submit_w_ccs(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, uint64_t *cm, uint64_t *ticket) {
    if (!h || !w_ccs || !cm || !ticket) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (w_len * h->L != h->n) return h->wrong_len(w_len * h->L);
    int st = h->bind();
    if (st || (st = h->matrix_ready())) return st;
    lat_ajtai::Slot &sl = h->slots[h->next_ticket % LAT_PIPELINE_DEPTH];
    if (sl.busy)
        return fail(LAT_E_INVALID_ARGUMENT, "pipeline full: lat_ajtai_wait(ticket " + std::to_string(sl.ticket) + ") first");
    const size_t in_bytes = w_len * ELEM_BYTES;  // fits: the slot was sized for ceil(n / L) elements at creation
    const unsigned long long tk = h->next_ticket;
    h->guard.timeout_ns = spin_timeout_ns();
    // Upload on the copy stream, followed by a copy of the ticket: the witness kernel polls that word instead of the
    // compute stream waiting for an event, so the compute stream is a pure chain of kernels -- witness, matrix-vector,
    // witness, ... -- that overlap through programmatic dependent launches (see witness_kernel / mac_kernel).
    *sl.h_ticket = tk;
    CK(cudaMemcpyAsync(sl.in.p, w_ccs, in_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(sl.ready.p, sl.h_ticket, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->copy_stream));
    // this step's kernel may start under the previous step's draining matrix-vector kernel: other witness buffer
    u64 *fxp = h->fx.as<u64>();
    const bool chained = h->mac_was_last && !h->profiling;
    if (chained && h->last_mac_src == h->fx.p) fxp = h->fx_alt.as<u64>();
    lat::launch_witness(sl.in.as<u64>(), w_len, (int)h->log2_B, (int)h->L, h->mont, false, h->f16.as<int16_t>(), nullptr, nullptr,
                        fxp, sl.flag.as<int>(), h->stream, chained, sl.ready.as<unsigned long long>(), tk, h->guard);
    CK(cudaGetLastError());
    h->has_resident = true;
    // the matrix-vector kernel reports straight into mapped host memory: commitment, overflow flag, then the ticket
    lat::MacReport rep;
    unsigned long long *done_dev = nullptr;
    if ((st = device_view(sl.h_cm, (void **)&rep.cm_host)) || (st = device_view(sl.h_flag, (void **)&rep.flag_host)) ||
        (st = device_view(const_cast<unsigned long long *>(sl.h_done), (void **)&done_dev)))
        return st;
    rep.flag_dev = sl.flag.as<int>();
    if (!h->has_peers) {
        rep.done_host = done_dev;
        rep.done_value = tk;
        if ((st = h->mac_fx(fxp, h->n, 1, sl.cm.as<u64>(), rep))) return st;
    } else {
        // sharded: the matrix-vector kernel only moves the overflow flag; the exchange kernel behind it (part of the
        // same kernel chain) sums the partial commitments of all ranks and reports the full one
        u64 *cm_map = rep.cm_host;
        rep.cm_host = nullptr;
        if ((st = h->mac_fx(fxp, h->n, 1, sl.cm.as<u64>(), rep))) return st;
        lat::launch_exchange(sl.cm.as<u64>(), (u64)h->kappa * LAT_RING_DEGREE, h->peer_rank, h->peer_world, h->peers,
                             h->peer_epoch++, sl.out.as<u64>(), h->stream, cm_map, done_dev, tk, h->guard);
        CK(cudaGetLastError());
    }
    sl.busy = true;
    sl.user_cm = cm;
    sl.ticket = tk;
    *ticket = h->next_ticket++;
    return LAT_OK;
}

int lat_ajtai_set_peers(lat_ajtai *h, int rank, int world, const uint64_t *recv_ptrs, const uint64_t *flag_ptrs,
                        uint64_t next_epoch) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    if (world <= 1 || !recv_ptrs || !flag_ptrs) {
        h->has_peers = false;
        return LAT_OK;
    }
    if (world > lat::MAX_PEERS || rank < 0 || rank >= world || next_epoch == 0)
        return fail(LAT_E_INVALID_ARGUMENT, "need world <= 16, 0 <= rank < world, next_epoch >= 1");
    for (int r = 0; r < world; ++r) {
        h->peers.recv[r] = reinterpret_cast<u64 *>(recv_ptrs[r]);
        h->peers.flags[r] = reinterpret_cast<u64 *>(flag_ptrs[r]);
    }
    h->peer_rank = rank;
    h->peer_world = world;
    h->peer_epoch = next_epoch;
    h->has_peers = true;
    return LAT_OK;
}

int lat_ajtai_wait(lat_ajtai *h, uint64_t ticket) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    lat_ajtai::Slot &sl = h->slots[ticket % LAT_PIPELINE_DEPTH];
    if (!sl.busy || sl.ticket != ticket) return fail(LAT_E_INVALID_ARGUMENT, "no such ticket in flight");
    CK(cudaSetDevice(h->device));  // not bind(): waiting enqueues nothing, the kernel chain stays intact
    sl.busy = false;
    // poll the ticket the last CTA publishes; look at the stream now and then so that a device fault cannot hang us
    for (unsigned spins = 0; *sl.h_done != ticket; ++spins) {
        if ((spins & 0x3ff) == 0x3ff) std::this_thread::yield();  // several ranks may share the host's cores
        if ((spins & 0xffff) == 0xffff) {
            cudaError_t e = cudaStreamQuery(h->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) return fail_cuda(e, "lat_ajtai_wait", __LINE__);
            if (e == cudaSuccess && *sl.h_done != ticket)
                return fail(LAT_E_CUDA, "stream idle but the step never reported completion");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (int st = spin_status(h->device)) return st;  // a bounded device-side wait gave up: no result to hand out
    memcpy(sl.user_cm, sl.h_cm, (size_t)h->kappa * ELEM_BYTES);
    if (*sl.h_flag)
        return fail(LAT_E_DIGIT_OVERFLOW, "a coefficient needs more digits than the decomposition padding allows");
    return LAT_OK;
}

// ---- decompose_witness + commit_witnesses ---------------------------------------------------------------------------
// f16 already holds the coefficients; produce planes (to caller buffers or internal), K-1 commits and y_0.
// planes_only: stop after the planes (the fold step commits both sides' planes in one launch of its own).
static int planes_core(lat_ajtai *h, const u64 *cm_dev, u64 *planes_coeff_dev, u64 *planes_f_dev, u64 *cms_dev,
                       const int16_t *src16 = nullptr, bool planes_only = false) {
    int st;
    if (!src16) src16 = h->f16.as<int16_t>();
    // the extended-layout planes of this side stay resident for lat_ajtai_fold_witness
    if ((st = h->ensure_planes())) return st;
    u64 *pfx = h->planes_k1(h->cur_side);
    h->side_ready[h->cur_side] = true;
    {
        if (!h->planes_lut.p) {  // once per handle: the subset-sum table of the planes' transform
            if ((st = h->planes_lut.ensure(lat::PLANES_LUT_WORDS * sizeof(u64)))) return st;
            lat::launch_planes_lut(h->mont, h->planes_lut.as<u64>(), h->stream);
        }
        lat::launch_planes(src16, h->n, (int)h->K, h->mont, h->planes_lut.as<u64>(), planes_f_dev, pfx, h->planes_k0(h->cur_side),
                           planes_coeff_dev, h->stream);
        CK(cudaGetLastError());
    }
    if (cms_dev && !planes_only) {
        if (h->K > 1) {
            st = h->mac_fx(pfx, h->n, h->K - 1, cms_dev + (size_t)h->kappa * LAT_RING_DEGREE,
                           lat::MacReport(), true);  // planes are in the Toom-3 form
            if (st) return st;
        }
        if (cm_dev) {  // column-sharded callers derive y_0 after the exchange (lat_commitment_y0_dev)
            lat::launch_y0(cm_dev, cms_dev, h->K, h->kappa, h->stream);
            CK(cudaGetLastError());
            h->last_op_was_mac = false;
        }
    }
    return LAT_OK;
}

int lat_ajtai_decompose_commit_dev(lat_ajtai *h, const uint64_t *f_coeff_dev, uint64_t n, const uint64_t *cm_dev,
                                   uint64_t *planes_coeff_dev, uint64_t *planes_f_dev, uint64_t *cms_dev) {
    if (!h || !f_coeff_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (n != h->n) return h->wrong_len(n);
    int st = h->bind();
    if (st) return st;
    if (cms_dev && (st = h->matrix_ready())) return st;
    lat::launch_pack_coeff((const u64 *)f_coeff_dev, n, h->mont, (int)h->K, h->f16.as<int16_t>(), h->flag.as<int>(),
                           h->stream);
    CK(cudaGetLastError());
    h->has_resident = true;
    return planes_core(h, (const u64 *)cm_dev, (u64 *)planes_coeff_dev, (u64 *)planes_f_dev, (u64 *)cms_dev);
}

static int planes_host(lat_ajtai *h, const uint64_t *cm, uint64_t *planes_coeff, uint64_t *planes_f, uint64_t *cms) {
    int st;
    size_t plane_bytes = (size_t)h->K * h->n * ELEM_BYTES, cm_bytes = (size_t)h->kappa * ELEM_BYTES;
    if (cms) {
        if (!cm) return fail(LAT_E_INVALID_ARGUMENT, "cm is required to derive cms[0]");
        if ((st = h->cms.ensure((size_t)h->K * cm_bytes))) return st;
        CK(cudaMemcpyAsync(h->cm_in.p, cm, cm_bytes, cudaMemcpyHostToDevice, h->stream));
    }
    if (planes_coeff && (st = h->planes_coeff.ensure(plane_bytes))) return st;
    if (planes_f && (st = h->planes.ensure(plane_bytes))) return st;
    st = planes_core(h, h->cm_in.as<u64>(), planes_coeff ? h->planes_coeff.as<u64>() : nullptr,
                     planes_f ? h->planes.as<u64>() : nullptr, cms ? h->cms.as<u64>() : nullptr);
    if (st) return st;
    if (cms) CK(cudaMemcpyAsync(cms, h->cms.p, (size_t)h->K * cm_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (planes_f) CK(cudaMemcpyAsync(planes_f, h->planes.p, plane_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (planes_coeff) CK(cudaMemcpyAsync(planes_coeff, h->planes_coeff.p, plane_bytes, cudaMemcpyDeviceToHost, h->stream));
    return h->finish();
}

int lat_ajtai_decompose_commit(lat_ajtai *h, const uint64_t *f_coeff, uint64_t n, const uint64_t *cm,
                               uint64_t *planes_coeff, uint64_t *planes_f, uint64_t *cms) {
    if (!h || !f_coeff) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (n != h->n) return h->wrong_len(n);
    int st = h->bind();
    if (st) return st;
    if (cms && (st = h->matrix_ready())) return st;
    size_t in_bytes = n * ELEM_BYTES;
    if ((st = h->in.ensure(in_bytes))) return st;
    CK(cudaMemcpyAsync(h->in.p, f_coeff, in_bytes, cudaMemcpyHostToDevice, h->stream));
    lat::launch_pack_coeff(h->in.as<u64>(), n, h->mont, (int)h->K, h->f16.as<int16_t>(), h->flag.as<int>(), h->stream);
    CK(cudaGetLastError());
    h->has_resident = true;
    return planes_host(h, cm, planes_coeff, planes_f, cms);
}

int lat_ajtai_decompose_commit_resident(lat_ajtai *h, const uint64_t *cm, uint64_t *planes_coeff, uint64_t *planes_f,
                                        uint64_t *cms) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    if (!h->has_resident) return fail(LAT_E_INVALID_ARGUMENT, "no resident witness: call lat_ajtai_witness_from_w_ccs first");
    int st = h->bind();
    if (st) return st;
    if (cms && (st = h->matrix_ready())) return st;
    // the resident digits are base-B limbs (|c| <= B/2 <= 2^14 < 2^K for the zkVM parameters); re-check the bound
    // only when it could fail
    if (h->log2_B > h->K) {
        return fail(LAT_E_INVALID_ARGUMENT, "resident limbs may exceed 2^K (log2_B > K): use lat_ajtai_decompose_commit");
    }
    return planes_host(h, cm, planes_coeff, planes_f, cms);
}

// ---- compute_f_0 + Witness::from_f --------------------------------------------------------------------------------------
int lat_ajtai_select_side(lat_ajtai *h, int side) {
    if (!h || (side != 0 && side != 1)) return fail(LAT_E_INVALID_ARGUMENT, "side must be 0 or 1");
    h->cur_side = side;
    return LAT_OK;
}

int lat_ajtai_fold_witness_dev(lat_ajtai *h, const uint64_t *rho_dev, uint64_t *f0_dev, uint64_t *f0_coeff_dev) {
    if (!h || !rho_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (!h->side_ready[0] || !h->side_ready[1])
        return fail(LAT_E_INVALID_ARGUMENT, "both sides must be decomposed first (lat_ajtai_select_side + lat_ajtai_decompose_commit)");
    int st = h->bind();
    if (st) return st;
    u64 *f0 = (u64 *)f0_dev;
    if (!f0) {
        if ((st = h->f0.ensure(h->n * ELEM_BYTES))) return st;
        f0 = h->f0.as<u64>();
    }
    const u64 *sides[2] = {h->planes_k1(0), h->planes_k1(1)}, *sides0[2] = {h->planes_k0(0), h->planes_k0(1)};
    lat::launch_fold(sides, sides0, 2, (int)h->K, h->n, (const u64 *)rho_dev, h->mont, f0, h->stream);
    CK(cudaGetLastError());
    if (f0_coeff_dev) {
        lat::launch_icrt(f0, (u64 *)f0_coeff_dev, h->n, h->stream);
        CK(cudaGetLastError());
    }
    return LAT_OK;
}

int lat_ajtai_fold_witness(lat_ajtai *h, const uint64_t *rho, uint64_t *f0, uint64_t *f0_coeff) {
    if (!h || !rho) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    int st = h->bind();
    if (st) return st;
    size_t n_bytes = h->n * ELEM_BYTES, rho_bytes = (size_t)2 * h->K * ELEM_BYTES;
    if ((st = h->rho.ensure(rho_bytes)) || (st = h->f0.ensure(n_bytes))) return st;
    if (f0_coeff && (st = h->in.ensure(n_bytes))) return st;
    CK(cudaMemcpyAsync(h->rho.p, rho, rho_bytes, cudaMemcpyHostToDevice, h->stream));
    st = lat_ajtai_fold_witness_dev(h, h->rho.as<uint64_t>(), h->f0.as<uint64_t>(), f0_coeff ? h->in.as<uint64_t>() : nullptr);
    if (st) return st;
    if (f0) CK(cudaMemcpyAsync(f0, h->f0.p, n_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (f0_coeff) CK(cudaMemcpyAsync(f0_coeff, h->in.p, n_bytes, cudaMemcpyDeviceToHost, h->stream));
    return h->finish();
}

// ---- the fold step as two blocking calls (zk_latticefold.rs:37-102) -------------------------------------------------------
// begin : Witness::from_w_ccs + commit of the step witness, then decompose_witness + commit_witnesses of BOTH
//         decompositions -- side 0 = the running accumulator (resident from the previous finish or from
//         lat_ajtai_set_accumulator), side 1 = the step witness -- chained on the device, one synchronisation.
// finish: compute_f_0 over the 2K resident planes with the host's challenges, Witness::from_f's iCRT, the folded
//         commitment cm_0 = sum rho_i cm_i, and the re-packing of f_0's coefficients as the next accumulator.
// Between the two the host runs its sumchecks (the challenges depend on all 2K commitments).
static int set_accumulator_dev(lat_ajtai *h, const u64 *f_coeff_dev) {
    int st = h->f16_acc.ensure(h->n * LAT_RING_DEGREE * sizeof(int16_t));
    if (st) return st;
    lat::launch_pack_coeff(f_coeff_dev, h->n, h->mont, (int)h->K, h->f16_acc.as<int16_t>(), h->flag.as<int>(), h->stream);
    CK(cudaGetLastError());
    h->acc_ready = true;
    return LAT_OK;
}

int lat_ajtai_set_accumulator(lat_ajtai *h, const uint64_t *f_coeff, uint64_t n, const uint64_t *cm_acc) {
    if (!h || !f_coeff) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (n != h->n) return h->wrong_len(n);
    int st = h->bind();
    if (st) return st;
    const size_t n_bytes = n * ELEM_BYTES, cm_bytes = (size_t)h->kappa * ELEM_BYTES;
    if ((st = h->in.ensure(n_bytes)) || (st = h->cm_acc.ensure(cm_bytes))) return st;
    CK(cudaMemcpyAsync(h->in.p, f_coeff, n_bytes, cudaMemcpyHostToDevice, h->stream));
    if ((st = set_accumulator_dev(h, h->in.as<u64>()))) return st;
    h->cm_acc_ready = false;
    if (cm_acc) {
        CK(cudaMemcpyAsync(h->cm_acc.p, cm_acc, cm_bytes, cudaMemcpyHostToDevice, h->stream));
        h->cm_acc_ready = true;
    }
    st = h->finish();
    if (st) h->acc_ready = false;
    return st;
}

int lat_ajtai_fold_step_begin(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, const uint64_t *cm_acc,
                              int16_t *f_coeff16, uint64_t *cm, uint64_t *cms) {
    if (!h || !w_ccs || !cm || !cms) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (w_len * h->L != h->n) return h->wrong_len(w_len * h->L);
    if (!h->acc_ready) return fail(LAT_E_INVALID_ARGUMENT, "no accumulator witness: call lat_ajtai_set_accumulator first");
    if (!cm_acc && !h->cm_acc_ready) return fail(LAT_E_INVALID_ARGUMENT, "no accumulator commitment resident: pass cm_acc");
    if (h->log2_B > h->K) return fail(LAT_E_INVALID_ARGUMENT, "step limbs may exceed 2^K (log2_B > K)");
    int st = h->bind();
    if (st || (st = h->matrix_ready())) return st;
    const size_t cm_bytes = (size_t)h->kappa * ELEM_BYTES, side_bytes = (size_t)h->K * cm_bytes;
    if ((st = h->cm_step.ensure(cm_bytes)) || (st = h->cm_acc.ensure(cm_bytes)) || (st = h->cms_side[0].ensure(side_bytes)) ||
        (st = h->cms_side[1].ensure(side_bytes)))
        return st;
    if (cm_acc) CK(cudaMemcpyAsync(h->cm_acc.p, cm_acc, cm_bytes, cudaMemcpyHostToDevice, h->stream));
    // The accumulator's side does not depend on this step's witness: its 15 planes go first, and w_ccs comes up on the
    // copy engine underneath them (every call ends in finish(), so h->in is free and copy_stream idle).  The 2 (K-1)
    // planes that get committed sit back to back in planes_all, so ONE launch commits both sides: 7 groups of 4 planes
    // share every matrix tile in L2 (one pass over HBM instead of four) and there is no 2-plane tail launch that would
    // stream the whole 5-word matrix for two planes.
    if ((st = h->in.ensure(w_len * ELEM_BYTES))) return st;
    CK(cudaEventRecord(h->work_done, h->stream));  // whatever the caller still has in flight on the handle's stream goes first
    CK(cudaStreamWaitEvent(h->copy_stream, h->work_done, 0));
    CK(cudaMemcpyAsync(h->in.p, w_ccs, w_len * ELEM_BYTES, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->copy_done[0], h->copy_stream));
    // side 0: the accumulator (acc, w_acc); side 1: the step witness (lin_cm_i, w_i)      zk_latticefold.rs:60-71
    const int side_before = h->cur_side;
    h->cur_side = 0;
    st = planes_core(h, nullptr, nullptr, nullptr, nullptr, h->f16_acc.as<int16_t>(), true);
    if (!st) {
        // the step witness and its commitment (ZKVM/main.rs:348-367)
        CK(cudaStreamWaitEvent(h->stream, h->copy_done[0], 0));
        st = witness_enqueue(h, w_ccs, w_len, false, nullptr, nullptr, h->cm_step.as<u64>(), f_coeff16 != nullptr, lat::MacReport(), true);
    }
    if (!st) {
        CK(cudaMemcpyAsync(cm, h->cm_step.p, cm_bytes, cudaMemcpyDeviceToHost, h->stream));
        if (f_coeff16)  // on copy_stream, beside the 14 matrix-vector products of side 1 (the digits are not modified by them)
            CK(cudaMemcpyAsync(f_coeff16, h->f16.p, h->n * LAT_RING_DEGREE * sizeof(int16_t), cudaMemcpyDeviceToHost, h->copy_stream));
        h->cur_side = 1;
        st = planes_core(h, nullptr, nullptr, nullptr, nullptr, h->f16.as<int16_t>(), true);
    }
    h->cur_side = side_before;
    if (st) return st;
    if (h->K > 1) {  // commit_witnesses of both sides (decomposition.rs:185-187), one launch
        const uint32_t per_side = h->K - 1;
        if ((st = h->cms.ensure((size_t)2 * per_side * cm_bytes))) return st;
        if ((st = h->mac_fx(h->planes_k1(0), h->n, 2 * per_side, h->cms.as<u64>(), lat::MacReport(), true))) return st;
        for (int s = 0; s < 2; ++s)  // behind each side's slot for y_0
            CK(cudaMemcpyAsync(h->cms_side[s].as<u64>() + (size_t)h->kappa * LAT_RING_DEGREE,
                               h->cms.as<u64>() + (size_t)s * per_side * h->kappa * LAT_RING_DEGREE, per_side * cm_bytes,
                               cudaMemcpyDeviceToDevice, h->stream));
    }
    // y_0 = cm - sum 2^k y_k for either side (decomposition.rs:189-197)
    lat::launch_y0(h->cm_acc.as<u64>(), h->cms_side[0].as<u64>(), h->K, h->kappa, h->stream);
    lat::launch_y0(h->cm_step.as<u64>(), h->cms_side[1].as<u64>(), h->K, h->kappa, h->stream);
    CK(cudaGetLastError());
    h->last_op_was_mac = false;
    CK(cudaMemcpyAsync(cms, h->cms_side[0].p, side_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(cms + (size_t)h->K * h->kappa * LAT_RING_DEGREE, h->cms_side[1].p, side_bytes, cudaMemcpyDeviceToHost,
                       h->stream));
    h->fold_pending = true;
    return h->finish();
}

int lat_ajtai_fold_step_finish(lat_ajtai *h, const uint64_t *rho, int16_t *f0_coeff16, uint64_t *f0, uint64_t *cm0,
                               uint64_t *w_ccs0) {
    if (!h || !rho) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (!h->fold_pending) return fail(LAT_E_INVALID_ARGUMENT, "lat_ajtai_fold_step_begin has not run");
    int st = h->bind();
    if (st) return st;
    const size_t n_bytes = h->n * ELEM_BYTES, rho_bytes = (size_t)2 * h->K * ELEM_BYTES, cm_bytes = (size_t)h->kappa * ELEM_BYTES;
    if ((st = h->rho.ensure(rho_bytes)) || (st = h->f0.ensure(n_bytes)) || (st = h->in.ensure(n_bytes))) return st;
    CK(cudaMemcpyAsync(h->rho.p, rho, rho_bytes, cudaMemcpyHostToDevice, h->stream));
    if (!h->side_ready[0] || !h->side_ready[1]) return fail(LAT_E_INVALID_ARGUMENT, "both sides must be decomposed first");
    if ((st = h->f16_acc.ensure(h->n * LAT_RING_DEGREE * sizeof(int16_t)))) return st;
    // f_0 = sum rho_i f_i (folding.rs:258-268), its coefficients (arith.rs:299-313), and those packed as the next
    // accumulator (they must stay below 2^K, the protocol's norm bound) -- in four ranges of elements when the caller wants
    // the digits back, so that the 4.7 MB download of a range runs under the fold of the next one
    {
        const u64 *sides[2] = {h->planes_k1(0), h->planes_k1(1)}, *sides0[2] = {h->planes_k0(0), h->planes_k0(1)};
        const int nch = (f0_coeff16 && h->n >= 4096) ? 4 : 1;
        for (int c = 0; c < nch; ++c) {
            const u64 e0 = h->n * c / nch, cnt = h->n * (c + 1) / nch - e0;
            lat::launch_fold(sides, sides0, 2, (int)h->K, h->n, h->rho.as<u64>(), h->mont, h->f0.as<u64>(), h->stream, e0, cnt);
            lat::launch_icrt(h->f0.as<u64>() + e0 * LAT_RING_DEGREE, h->in.as<u64>() + e0 * LAT_RING_DEGREE, cnt, h->stream);
            lat::launch_pack_coeff(h->in.as<u64>() + e0 * LAT_RING_DEGREE, cnt, h->mont, (int)h->K,
                                   h->f16_acc.as<int16_t>() + e0 * LAT_RING_DEGREE, h->flag.as<int>(), h->stream);
            CK(cudaGetLastError());
            if (f0_coeff16) {
                CK(cudaEventRecord(h->copy_done[c], h->stream));
                CK(cudaStreamWaitEvent(h->copy_stream, h->copy_done[c], 0));
                CK(cudaMemcpyAsync(f0_coeff16 + e0 * LAT_RING_DEGREE, h->f16_acc.as<int16_t>() + e0 * LAT_RING_DEGREE,
                                   cnt * LAT_RING_DEGREE * sizeof(int16_t), cudaMemcpyDeviceToHost, h->copy_stream));
                h->copy_stream_busy = true;
            }
        }
        h->acc_ready = true;
    }
    // cm_0 = sum rho_i cm_i (folding/utils.rs:466-472): the next step's accumulator commitment, kept resident
    lat::launch_lincomb(h->rho.as<u64>(), h->cms_side[0].as<u64>(), h->cms_side[1].as<u64>(), (int)h->K, h->kappa, h->mont,
                        h->cm_acc.as<u64>(), h->stream);
    CK(cudaGetLastError());
    h->cm_acc_ready = true;
    if (f0) CK(cudaMemcpyAsync(f0, h->f0.p, n_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (cm0) CK(cudaMemcpyAsync(cm0, h->cm_acc.p, cm_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (w_ccs0) {  // Witness::from_f rebuilds w_ccs = gadget_recompose(f_0) (arith.rs:305): n / L elements, CRT form
        const u64 count = h->n / h->L;
        // h->in held f_0's coefficients; they are packed into f16_acc by now (stream order), so it is free again
        lat::launch_recompose(h->f0.as<u64>(), count, (int)h->log2_B, (int)h->L, h->in.as<u64>(), h->stream);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(w_ccs0, h->in.p, count * ELEM_BYTES, cudaMemcpyDeviceToHost, h->stream));
    }
    h->fold_pending = false;
    st = h->finish();
    if (st) h->acc_ready = false;  // e.g. LAT_E_DIGIT_OVERFLOW: the folded witness broke the norm bound
    return st;
}

// Witness::get_fhat (LF/arith.rs:273-297) of a resident witness, on the device: tau = 3 tables of n ring elements, table j,
// element i, slot s = (coefficient 8j + s of f_coeff[i], 0, 0).  For MLE code that runs on the GPU; a host gets the same
// tables by re-laying out the int16 digits itself (scheme.get_fhat_from_digits / ajtai.hpp), which is 12x less PCIe.
int lat_ajtai_get_fhat_dev(lat_ajtai *h, int which, uint64_t *fhat_dev) {
    if (!h || !fhat_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (which != 0 && which != 1) return fail(LAT_E_INVALID_ARGUMENT, "which: 0 = current witness, 1 = accumulator");
    if (which == 0 ? !h->has_resident : !h->acc_ready) return fail(LAT_E_INVALID_ARGUMENT, "no such resident witness");
    int st = h->bind();
    if (st) return st;
    lat::launch_fhat((which == 0 ? h->f16 : h->f16_acc).as<int16_t>(), h->n, h->mont, (u64 *)fhat_dev, h->stream);
    CK(cudaGetLastError());
    return LAT_OK;
}

int lat_ring_gadget_recompose(const uint64_t *f, uint64_t count, uint32_t log2_b, uint32_t L, uint64_t *out, int repr,
                              int device) {
    (void)repr;  // a scalar multiple and sums: the same in either representation
    if (count == 0) return LAT_OK;
    if (!f || !out) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (log2_b < 1 || log2_b > 62 || L < 1) return fail(LAT_E_INVALID_ARGUMENT, "need 1<=log2_b<=62, L>=1");
    ScratchLock sc;
    int st = sc.open(device, count * L * ELEM_BYTES, count * ELEM_BYTES);
    if (st) return st;
    CK(cudaMemcpyAsync(sc.at<u64>(0), f, count * L * ELEM_BYTES, cudaMemcpyHostToDevice, sc.stream()));
    lat::launch_recompose(sc.at<u64>(0), count, (int)log2_b, (int)L, sc.at<u64>(1), sc.stream());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, sc.at<u64>(1), count * ELEM_BYTES, cudaMemcpyDeviceToHost, sc.stream()));
    CK(cudaStreamSynchronize(sc.stream()));
    return LAT_OK;
}

// ---- standalone transforms ---------------------------------------------------------------------------------------
int lat_ring_crt_dev(const uint64_t *coeff_dev, uint64_t count, uint64_t *ntt_dev, void *cuda_stream) {
    if (count && (!coeff_dev || !ntt_dev)) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    lat::launch_crt((const u64 *)coeff_dev, (u64 *)ntt_dev, count, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return LAT_OK;
}
int lat_ring_icrt_dev(const uint64_t *ntt_dev, uint64_t count, uint64_t *coeff_dev, void *cuda_stream) {
    if (count && (!ntt_dev || !coeff_dev)) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    lat::launch_icrt((const u64 *)ntt_dev, (u64 *)coeff_dev, count, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return LAT_OK;
}

static int ring_host(const uint64_t *in, uint64_t count, uint64_t *out, int device, bool inverse) {
    if (count == 0) return LAT_OK;
    if (!in || !out) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    ScratchLock sc;
    int st = sc.open(device, count * ELEM_BYTES);
    if (st) return st;
    CK(cudaMemcpyAsync(sc.at<u64>(0), in, count * ELEM_BYTES, cudaMemcpyHostToDevice, sc.stream()));
    if (inverse) lat::launch_icrt(sc.at<u64>(0), sc.at<u64>(0), count, sc.stream());  // in place
    else lat::launch_crt(sc.at<u64>(0), sc.at<u64>(0), count, sc.stream());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, sc.at<u64>(0), count * ELEM_BYTES, cudaMemcpyDeviceToHost, sc.stream()));
    CK(cudaStreamSynchronize(sc.stream()));
    return LAT_OK;
}
int lat_ring_crt(const uint64_t *coeff, uint64_t count, uint64_t *ntt, int device) {
    return ring_host(coeff, count, ntt, device, false);
}
int lat_ring_icrt(const uint64_t *ntt, uint64_t count, uint64_t *coeff, int device) {
    return ring_host(ntt, count, coeff, device, true);
}

int lat_ring_gadget_decompose(const uint64_t *in, uint64_t count, uint32_t log2_b, uint32_t L, uint64_t *out, int repr,
                              int device) {
    if (count == 0) return LAT_OK;
    if (!in || !out) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (log2_b < 1 || log2_b > 15 || L < 1 || L > 32) return fail(LAT_E_INVALID_ARGUMENT, "need 1<=log2_b<=15, 1<=L<=32");
    ScratchLock sc;
    int st = sc.open(device, count * ELEM_BYTES, count * L * LAT_RING_DEGREE * sizeof(int16_t), count * L * ELEM_BYTES, sizeof(int));
    if (st) return st;
    CK(cudaMemsetAsync(sc.at<int>(3), 0, sizeof(int), sc.stream()));
    CK(cudaMemcpyAsync(sc.at<u64>(0), in, count * ELEM_BYTES, cudaMemcpyHostToDevice, sc.stream()));
    lat::launch_witness(sc.at<u64>(0), count, (int)log2_b, (int)L, repr == LAT_REPR_MONTGOMERY, true, sc.at<int16_t>(1),
                        sc.at<u64>(2), nullptr, nullptr, sc.at<int>(3), sc.stream());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, sc.at<u64>(2), count * L * ELEM_BYTES, cudaMemcpyDeviceToHost, sc.stream()));
    CK(cudaMemcpyAsync(sc.s->h_flag, sc.at<int>(3), sizeof(int), cudaMemcpyDeviceToHost, sc.stream()));
    CK(cudaStreamSynchronize(sc.stream()));
    if (*sc.s->h_flag) return fail(LAT_E_DIGIT_OVERFLOW, "a coefficient needs more than L digits");
    return LAT_OK;
}

// ---- standalone power-of-two negacyclic NTT (absent from the reference) ----------------------------------------
int lat_ntt_negacyclic_dev(const uint64_t *in_dev, uint64_t batch, uint32_t log2_d, int inverse, uint64_t *out_dev,
                           void *cuda_stream) {
    if (batch == 0) return LAT_OK;
    if (!in_dev || !out_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (log2_d < 1 || log2_d > LAT_NTT_MAX_LOG2_D) return fail(LAT_E_INVALID_ARGUMENT, "need 1 <= log2_d <= 14");
    int e = lat::launch_ntt_pow2((const u64 *)in_dev, (u64 *)out_dev, batch, log2_d, inverse != 0, (cudaStream_t)cuda_stream);
    if (e) return fail_cuda((cudaError_t)e, "lat_ntt_negacyclic", __LINE__);
    return LAT_OK;
}

int lat_ntt_negacyclic(const uint64_t *in, uint64_t batch, uint32_t log2_d, int inverse, uint64_t *out, int device) {
    if (batch == 0) return LAT_OK;
    if (!in || !out) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (log2_d < 1 || log2_d > LAT_NTT_MAX_LOG2_D) return fail(LAT_E_INVALID_ARGUMENT, "need 1 <= log2_d <= 14");
    const size_t bytes = (size_t)(batch << log2_d) * sizeof(uint64_t);
    ScratchLock sc;
    int st = sc.open(device, bytes);
    if (st) return st;
    CK(cudaMemcpyAsync(sc.at<u64>(0), in, bytes, cudaMemcpyHostToDevice, sc.stream()));
    if ((st = lat_ntt_negacyclic_dev(sc.at<uint64_t>(0), batch, log2_d, inverse, sc.at<uint64_t>(0), sc.stream()))) return st;
    CK(cudaMemcpyAsync(out, sc.at<u64>(0), bytes, cudaMemcpyDeviceToHost, sc.stream()));
    CK(cudaStreamSynchronize(sc.stream()));
    return LAT_OK;
}

int lat_ajtai_set_step_overlap(lat_ajtai *h, int enabled) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    h->step_overlap = enabled != 0;
    return LAT_OK;
}

int lat_ajtai_set_profiling(lat_ajtai *h, int enabled) {
    if (!h) return fail(LAT_E_INVALID_ARGUMENT, "NULL handle");
    int st = h->bind();
    if (st) return st;
    if (enabled && h->ev0.empty()) {
        h->ev0.resize(lat_ajtai::EV_POOL);
        h->ev1.resize(lat_ajtai::EV_POOL);
        h->ev_pending.assign(lat_ajtai::EV_POOL, 0);
        for (int i = 0; i < lat_ajtai::EV_POOL; ++i) {
            CK(cudaEventCreate(&h->ev0[i]));
            CK(cudaEventCreate(&h->ev1[i]));
        }
    }
    h->profiling = enabled != 0;
    return LAT_OK;
}
int lat_ajtai_mac_profile(lat_ajtai *h, double *sum_ms, uint64_t *launches) {
    if (!h || !sum_ms || !launches) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    int st = h->bind();
    if (st) return st;
    for (size_t i = 0; i < h->ev_pending.size(); ++i)
        if ((st = h->drain_slot((int)i))) return st;
    *sum_ms = h->prof_sum_ms;
    *launches = h->prof_count;
    h->prof_sum_ms = 0.0;
    h->prof_count = 0;
    return LAT_OK;
}

int lat_commitment_y0_dev(const uint64_t *cm_dev, uint64_t *cms_dev, uint32_t K, uint32_t kappa, void *cuda_stream) {
    if (!cm_dev || !cms_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (K < 1 || kappa < 1) return fail(LAT_E_INVALID_ARGUMENT, "need K >= 1, kappa >= 1");
    lat::launch_y0((const u64 *)cm_dev, (u64 *)cms_dev, K, kappa, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return LAT_OK;
}

int lat_commitment_sum_dev(const uint64_t *parts_dev, uint32_t count, uint64_t words, uint64_t *out_dev,
                           void *cuda_stream) {
    if (words && (!parts_dev || !out_dev)) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    lat::launch_commitment_sum((const u64 *)parts_dev, count, words, (u64 *)out_dev, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    return LAT_OK;
}
int lat_commitment_exchange_report_dev(const uint64_t *partial_dev, uint64_t words, int rank, int world,
                                       const uint64_t *recv_ptrs, const uint64_t *flag_ptrs, uint64_t epoch,
                                       uint64_t *out_dev, uint64_t *cm_host, uint64_t *done_host, uint64_t done_value,
                                       void *cuda_stream) {
    if (!partial_dev || !recv_ptrs || !flag_ptrs || !out_dev) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    if (world < 1 || world > lat::MAX_PEERS || rank < 0 || rank >= world || epoch == 0)
        return fail(LAT_E_INVALID_ARGUMENT, "need 1 <= world <= 16, 0 <= rank < world, epoch >= 1");
    lat::PeerPtrs peers{};
    for (int r = 0; r < world; ++r) {
        peers.recv[r] = reinterpret_cast<u64 *>(recv_ptrs[r]);
        peers.flags[r] = reinterpret_cast<u64 *>(flag_ptrs[r]);
    }
    u64 *cm_map = nullptr;
    unsigned long long *done_map = nullptr;
    if (cm_host) CK(cudaHostGetDevicePointer((void **)&cm_map, cm_host, 0));
    if (done_host) CK(cudaHostGetDevicePointer((void **)&done_map, done_host, 0));
    lat::SpinGuard guard;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (int st = spin_guard(dev, guard)) return st;
    lat::launch_exchange((const u64 *)partial_dev, words, rank, world, peers, epoch, (u64 *)out_dev, (cudaStream_t)cuda_stream,
                         cm_map, done_map, done_value, guard);
    CK(cudaGetLastError());
    return LAT_OK;
}

int lat_commitment_exchange_dev(const uint64_t *partial_dev, uint64_t words, int rank, int world,
                                const uint64_t *recv_ptrs, const uint64_t *flag_ptrs, uint64_t epoch,
                                uint64_t *out_dev, void *cuda_stream) {
    return lat_commitment_exchange_report_dev(partial_dev, words, rank, world, recv_ptrs, flag_ptrs, epoch, out_dev, nullptr,
                                              nullptr, 0, cuda_stream);
}

int lat_commitment_sum(const uint64_t *parts, uint32_t count, uint64_t words, uint64_t *out, int device) {
    if (words == 0) return LAT_OK;
    if (!parts || !out) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    ScratchLock sc;
    int st = sc.open(device, (size_t)count * words * 8, words * 8);
    if (st) return st;
    CK(cudaMemcpyAsync(sc.at<u64>(0), parts, (size_t)count * words * 8, cudaMemcpyHostToDevice, sc.stream()));
    lat::launch_commitment_sum(sc.at<u64>(0), count, words, sc.at<u64>(1), sc.stream());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, sc.at<u64>(1), words * 8, cudaMemcpyDeviceToHost, sc.stream()));
    CK(cudaStreamSynchronize(sc.stream()));
    return LAT_OK;
}

int lat_set_spin_timeout_ms(uint64_t ms) {
    g_spin_timeout_ns.store(ms >= (~0ull - 1) / 1000000ull ? 0ull : ms * 1000000ull, std::memory_order_relaxed);
    return LAT_OK;
}
int lat_device_wait_status(int device, uint64_t *code) {
    if (device < 0 || device >= MAX_DEVICES) return fail(LAT_E_INVALID_ARGUMENT, "device ordinal out of range");
    return spin_status(device, code);
}

int lat_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(LAT_E_INVALID_ARGUMENT, "NULL argument");
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return LAT_OK;
}
void lat_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
