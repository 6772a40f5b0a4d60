// Multi-GPU size selection inside the library: one process per GPU, the ranks' handles joined by an
// NCCL communicator of their own, every exchange queued on the handle's stream - no host framework between the steps.
//
//   sampler.go:65-106 on N ranks = all-gather of the per-length pool totals T_r[l] (the only data north_star lets cross
//   NVLink besides the routed draws; the partial target histograms ride along), class-partitioned draw (class l by rank
//   l % N), one grouped send/recv that routes every drawn copy rank to the rank owning the copy, apply, expand.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 - the instance the process already loaded, if any), so the library
// has no link-time dependency on it and a single-GPU host never needs it.
#include <dlfcn.h>
#include <nccl.h>
#include "handle.cuh"

using namespace rl_api;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("RLSIM_NCCL_LIB");
        if (env) api.lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already uses
        if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(api.lib, name);
            if (!p && api.err.empty()) api.err = std::string("libnccl: missing symbol ") + name;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}

#define NK(h, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { (h)->err = std::string(#call) + ": " + nccl_api()->GetErrorString(r__); return RLSIM_E_CUDA; } } while (0)

int ensure_pin(rlsim_handle *h, size_t n) {
    if (h->x_pin_n >= n) return 0;
    if (h->x_pin) cudaFreeHost(h->x_pin);
    h->x_pin = nullptr; h->x_pin_n = 0;
    CK(h, cudaHostAlloc((void **)&h->x_pin, (n + n / 4 + 64) * 8, cudaHostAllocDefault));
    h->x_pin_n = n + n / 4 + 64;
    return 0;
}

// all-gather `n` u64 per rank (device row d_row) -> host matrix [N][n] in h->x_pin (valid until the next gather)
int gather_rows(rlsim_handle *h, const uint64_t *d_row, size_t n) {
    NcclApi *A = nccl_api();
    const size_t N = (size_t)h->comm_n;
    int rc;
    CK(h, h->d_x_recv.ensure(N * n));
    if ((rc = ensure_pin(h, N * n))) return rc;
    NK(h, A->AllGather(d_row, h->d_x_recv.p, n, ncclUint64, h->comm, h->st));
    CK(h, cudaMemcpyAsync(h->x_pin, h->d_x_recv.p, N * n * 8, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    return 0;
}

}  // namespace

namespace rl_api {
void comm_destroy(rlsim_handle *h) {
    if (h->comm) { nccl_api()->CommDestroy(h->comm); h->comm = nullptr; }
    if (h->x_pin) { cudaFreeHost(h->x_pin); h->x_pin = nullptr; h->x_pin_n = 0; }
}
}  // namespace rl_api

extern "C" {

int rlsim_comm_unique_id(uint8_t *id_out) {
    NcclApi *A = nccl_api();
    if (!A->err.empty() || !A->GetUniqueId) return RLSIM_E_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == RLSIM_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (A->GetUniqueId(&id) != ncclSuccess) return RLSIM_E_CUDA;
    memcpy(id_out, &id, sizeof id);
    return 0;
}

int rlsim_comm_init(rlsim_handle *h, const uint8_t *id_bytes, int nranks, int rank) {
    NcclApi *A = nccl_api();
    if (!A->err.empty()) FAIL(h, RLSIM_E_UNSUPPORTED, A->err);
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) FAIL(h, RLSIM_E_ARG, "invalid rank / number of ranks (1..64)");
    cudaSetDevice(h->device);
    comm_destroy(h);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    NK(h, A->CommInitRank(&h->comm, nranks, id, rank));
    h->comm_n = nranks; h->comm_rank = rank;
    return 0;
}

// pure host arithmetic of the record routing (CPU-testable): counts[s * n + d] = records rank s holds for rank d.
// send_off[n+1] = where this rank's groups start in its send buffer; recv_off[n+1] = where each sender's records land.
int rlsim_exchange_plan(const uint64_t *counts, uint32_t n, uint32_t rank, uint64_t *send_off, uint64_t *recv_off) {
    if (!n || rank >= n) return RLSIM_E_ARG;
    send_off[0] = recv_off[0] = 0;
    for (uint32_t r = 0; r < n; r++) {
        send_off[r + 1] = send_off[r] + counts[(size_t)rank * n + r];
        recv_off[r + 1] = recv_off[r] + counts[(size_t)r * n + rank];
    }
    return 0;
}

// Multi-GPU: sum the ranks' partial target histograms (rlsim_sample_target_range) now - a raw (-j) target shapes the
// library (range, mean), so its sum is needed before the transcripts are added.
int rlsim_comm_sum_target(rlsim_handle *h) {
    if (!h->comm) FAIL(h, RLSIM_E_ARG, "rlsim_comm_init first");
    if (!h->target_pending) FAIL(h, RLSIM_E_ARG, "no partial target histogram (rlsim_sample_target_range first)");
    cudaSetDevice(h->device);
    const uint32_t hn = h->target_hist_n;
    CK(h, cudaEventRecord(h->ev_aux, h->st_aux));
    CK(h, cudaStreamWaitEvent(h->st, h->ev_aux, 0));
    int rc = gather_rows(h, reinterpret_cast<const uint64_t *>(h->d_hist_target.p), hn);
    if (rc) return rc;
    std::vector<uint64_t> sum(hn, 0);
    for (int r = 0; r < h->comm_n; r++)
        for (uint32_t l = 0; l < hn; l++) sum[l] += h->x_pin[(size_t)r * hn + l];
    set_target_from_hist(h, sum.data(), hn);
    return 0;
}

// Sampler.SampleFragments (sampler.go:108-154) on the communicator's ranks, in one call.
int rlsim_sample_distributed(rlsim_handle *h) {
    if (!h->comm) FAIL(h, RLSIM_E_ARG, "rlsim_comm_init first");
    if (!h->amplified) FAIL(h, RLSIM_E_ARG, "rlsim_pcr first");
    NcclApi *A = nccl_api();
    cudaSetDevice(h->device);
    int rc;
    const uint32_t N = (uint32_t)h->comm_n, me = (uint32_t)h->comm_rank, n_len = h->n_len;
    const bool partial = h->target_pending && !h->have_target;  // every rank drew a slice of the target lengths
    if (!partial && !h->have_target) FAIL(h, RLSIM_E_ARG, "no target histogram");
    const uint32_t hn = partial ? h->target_hist_n : 0;
    const size_t row = (size_t)n_len + hn;
    // ---- 1. [per-length pool totals | partial target histogram] of every rank
    CK(h, h->d_x_send.ensure(row + 64));
    CK(h, cudaMemcpyAsync(h->d_x_send.p, h->d_totals.p, (size_t)n_len * 8, cudaMemcpyDeviceToDevice, h->st));
    if (partial) {
        CK(h, cudaEventRecord(h->ev_aux, h->st_aux));
        CK(h, cudaStreamWaitEvent(h->st, h->ev_aux, 0));
        CK(h, cudaMemcpyAsync(h->d_x_send.p + n_len, h->d_hist_target.p, (size_t)hn * 8, cudaMemcpyDeviceToDevice, h->st));
    }
    if ((rc = gather_rows(h, h->d_x_send.p, row))) return rc;
    std::vector<uint64_t> all_tot((size_t)N * n_len);
    h->h_global_totals.assign(n_len, 0);
    for (uint32_t r = 0; r < N; r++)
        for (uint32_t l = 0; l < n_len; l++) {
            const uint64_t v = h->x_pin[(size_t)r * row + l];
            all_tot[(size_t)r * n_len + l] = v;
            const uint64_t s = h->h_global_totals[l] + v;
            if (s < v) FAIL(h, RLSIM_E_POOL_OVERFLOW, "Integer overflow when updating total fragment count!");  // fragstats.go:88-94
            h->h_global_totals[l] = s;
        }
    if (partial) {
        std::vector<uint64_t> sum(hn, 0);
        for (uint32_t r = 0; r < N; r++)
            for (uint32_t l = 0; l < hn; l++) sum[l] += h->x_pin[(size_t)r * row + n_len + l];
        set_target_from_hist(h, sum.data(), hn);
    }
    // 8-byte records (length << 40 | copy rank) whenever every class total fits 40 bits; 16-byte ones otherwise
    bool rec8 = n_len <= (1u << 24);
    for (uint32_t l = 0; l < n_len && rec8; l++) rec8 = h->h_global_totals[l] < (1ull << 40);
    h->rec8 = rec8;
    const size_t u64_per_rec = rec8 ? 1 : 2;
    struct Rec8Off { rlsim_handle *h; ~Rec8Off() { h->rec8 = false; } } rec8_off{h};  // the C-ABI record form stays 16 bytes
    // ---- 2. draw this rank's classes; records grouped by owner rank in the send buffer
    uint64_t counts[64] = {0}, seg[64], needed = 0;
    uint64_t cap = std::max<uint64_t>(h->d_rec_send.n, h->req_frags / N + h->req_frags / (8 * N) + 4096);
    for (int attempt = 0;; attempt++) {
        if (attempt > 3) FAIL(h, RLSIM_E_NOMEM, "selection record buffer kept overflowing");
        CK(h, h->d_rec_send.ensure(cap));
        if ((rc = select_partitioned(h, all_tot.data(), N, me, n_len, h->d_rec_send.p, h->d_rec_send.n, counts, &needed, false, seg))) return rc;
        if (needed <= h->d_rec_send.n) break;
        cap = needed + needed / 8 + 4096;
    }
    // ---- 3. who sends how much to whom
    if ((rc = ensure_pin(h, (size_t)N * N + 64))) return rc;
    CK(h, h->d_x_send.ensure(std::max<size_t>(row, N) + 64));
    uint64_t *cnt_pin = h->x_pin + h->x_pin_n - 64;  // tail of the pinned zone: the gather lands at its head
    for (uint32_t r = 0; r < N; r++) cnt_pin[r] = counts[r];
    CK(h, cudaMemcpyAsync(h->d_x_send.p, cnt_pin, (size_t)N * 8, cudaMemcpyHostToDevice, h->st));
    if ((rc = gather_rows(h, h->d_x_send.p, N))) return rc;
    std::vector<uint64_t> send_off(N + 1), recv_off(N + 1);
    rlsim_exchange_plan(h->x_pin, N, me, send_off.data(), recv_off.data());
    if (seg[0] != ~0ull)  // one-sweep emit: the owner groups sit at their pre-sized offsets, not back to back
        for (uint32_t r = 0; r < N; r++) send_off[r] = seg[r];
    std::vector<uint64_t> from(N);
    for (uint32_t r = 0; r < N; r++) from[r] = h->x_pin[(size_t)r * N + me];
    const uint64_t total_recv = recv_off[N];
    CK(h, h->d_rec_recv.ensure(total_recv + 1));
    // ---- 4. one grouped exchange over NVLink on the handle's stream
    {
        StageTimer tm(h, RLSIM_T_EXCHANGE, 1);
        uint64_t *sendp = reinterpret_cast<uint64_t *>(h->d_rec_send.p), *recvp = reinterpret_cast<uint64_t *>(h->d_rec_recv.p);
        NK(h, A->GroupStart());
        for (uint32_t r = 0; r < N; r++) {
            if (r == me) continue;
            if (counts[r]) NK(h, A->Send(sendp + send_off[r] * u64_per_rec, (size_t)counts[r] * u64_per_rec, ncclUint64, (int)r, h->comm, h->st));
            if (from[r]) NK(h, A->Recv(recvp + recv_off[r] * u64_per_rec, (size_t)from[r] * u64_per_rec, ncclUint64, (int)r, h->comm, h->st));
        }
        NK(h, A->GroupEnd());
        if (counts[me])
            CK(h, cudaMemcpyAsync(recvp + recv_off[me] * u64_per_rec, sendp + send_off[me] * u64_per_rec, (size_t)counts[me] * 8 * u64_per_rec,
                                  cudaMemcpyDeviceToDevice, h->st));
        tm.stop();
    }
    // ---- 5. owners map the received copy ranks to their species; records out
    if ((rc = apply_records(h, h->d_rec_recv.p, total_recv))) return rc;
    return finish_sample(h);
}

}  // extern "C"
