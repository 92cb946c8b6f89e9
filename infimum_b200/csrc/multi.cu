// Multi-GPU poll-tree merge and batch hashing inside the library, for hosts
// that are ONE process driving several GPUs (the Rust shim of INTEGRATION.md
// cannot use torch.distributed).  Same sharding as infimum_b200/sharded.py:
// the logical leaf array is cut at a shard level into whole subtrees, each
// device reduces its contiguous run, the subtree roots are exchanged with ONE
// ncclAllGather over NVLink (group call over all devices, libnccl.so.2 loaded
// with dlopen so that single-GPU users carry no NCCL dependency), and device 0
// finishes the top levels.  Everything is enqueued asynchronously from one host
// thread on per-device streams; the only synchronisation is the final root read.
#include "infimum_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <pthread.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// internal entry points of capi.cu
extern "C" int inf_internal_stream(inf_ctx* ctx, void** stream);
extern "C" int inf_internal_grow_io(inf_ctx* ctx, int which, size_t bytes, void** ptr);
extern "C" int inf_internal_tree_reduce_host(inf_ctx* ctx, uint32_t arity, uint32_t n_levels, uint64_t shift,
                                             const uint8_t* h_in, uint64_t n_in, void* d_out, uint64_t* n_out);
extern "C" int inf_internal_drain(inf_ctx* ctx);

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::mutex mu;
    bool load() {
        std::lock_guard<std::mutex> lock(mu);       // inf_multi_init may be called from several threads
        if (handle) return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd;
    }
};
Nccl g_nccl;

uint64_t pow_sat(uint64_t a, uint32_t e) {
    unsigned __int128 r = 1;
    for (uint32_t i = 0; i < e; i++) {
        r *= a;
        if (r > (unsigned __int128)UINT64_MAX) return UINT64_MAX;
    }
    return (uint64_t)r;
}

// The calls below hop between devices; leave the caller's current device as it was.
struct DeviceRestore {
    int prev = -1;
    DeviceRestore() {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    }
    ~DeviceRestore() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct inf_multi {
    std::vector<int> devices;
    std::vector<inf_ctx*> ctx;
    std::vector<ncclComm_t> comms;       // empty when the gather runs over peer copies
    std::vector<void*> send, recv;       // per device gather buffers
    size_t gather_bytes = 0;
    bool use_nccl = false;
    std::string last_error;
};

extern "C" {

int inf_multi_init(const int* devices, int n_devices, uint32_t flags, inf_multi** out) {
    if (!devices || !out || n_devices < 1) return INF_ERR_NULL_POINTER;
    *out = nullptr;
    inf_multi* m = new inf_multi();
    m->devices.assign(devices, devices + n_devices);
    for (int d : m->devices) {
        inf_ctx* c = nullptr;
        int rc = inf_init(d, &c);
        if (rc) {
            inf_multi_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
    }
    m->use_nccl = n_devices > 1 && !(flags & INF_MULTI_PEER_COPY);
    if (m->use_nccl) {
        if (!g_nccl.load()) {
            inf_multi_destroy(m);
            return INF_ERR_NCCL;
        }
        m->comms.resize(n_devices);
        if (g_nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data()) != 0) {
            m->comms.clear();
            inf_multi_destroy(m);
            return INF_ERR_NCCL;
        }
    }
    m->send.assign(n_devices, nullptr);
    m->recv.assign(n_devices, nullptr);
    *out = m;
    return INF_OK;
}

void inf_multi_destroy(inf_multi* m) {
    if (!m) return;
    DeviceRestore restore;
    for (size_t i = 0; i < m->comms.size(); i++)
        if (m->comms[i]) g_nccl.CommDestroy(m->comms[i]);
    for (size_t i = 0; i < m->ctx.size(); i++) {
        cudaSetDevice(m->devices[i]);
        if (i < m->send.size() && m->send[i]) cudaFree(m->send[i]);
        if (i < m->recv.size() && m->recv[i]) cudaFree(m->recv[i]);
        inf_destroy(m->ctx[i]);
    }
    delete m;
}

int inf_multi_device_count(const inf_multi* m) { return m ? (int)m->devices.size() : 0; }

int inf_multi_tree_merge(inf_multi* m, uint32_t arity, uint32_t full_depth, int prepend_blank_leaf,
                         int to_depth, const uint8_t* leaves, uint64_t n_leaves, uint8_t root[32],
                         uint32_t* insert_depth, uint32_t* root_depth, int* has_root) {
    if (!m) return INF_ERR_NULL_POINTER;
    if (n_leaves && !leaves) return INF_ERR_NULL_POINTER;
    if (arity != 2 && arity != 5) return INF_ERR_BAD_ARITY;
    if (full_depth > 32) return INF_ERR_BAD_DEPTH;
    const int G = (int)m->devices.size();
    DeviceRestore restore;
    const uint64_t shift = prepend_blank_leaf ? 1 : 0, n_total = n_leaves + shift;
    const uint64_t cap = pow_sat(arity, full_depth);
    if (insert_depth) *insert_depth = 0;
    if (root_depth) *root_depth = 0;
    if (has_root) *has_root = 0;
    if (n_total > cap) return INF_ERR_TREE_ALREADY_FULL;
    if (n_total == 0) return INF_OK;
    uint32_t idepth = 0;
    while (idepth < full_depth && pow_sat(arity, idepth + 1) <= n_total) idepth++;
    const bool completed = n_total == cap;
    uint32_t rdepth = 0;
    if (to_depth || completed) rdepth = full_depth;
    else while (pow_sat(arity, rdepth) < n_total) rdepth++;
    if (insert_depth) *insert_depth = idepth;
    if (root_depth) *root_depth = rdepth;

    // shard level: at least 64 non-empty subtrees per device, so that the loads of
    // the devices, which differ by up to one subtree, agree within 1.6 %
    uint32_t k = 0;
    while (k + 1 <= rdepth && (n_total + pow_sat(arity, k + 1) - 1) / pow_sat(arity, k + 1) >= (uint64_t)64 * G) k++;
    const uint64_t w = pow_sat(arity, k);
    const uint64_t n_sub = (n_total + w - 1) / w;
    uint64_t width = 0;
    std::vector<uint64_t> s0(G), s1(G);
    for (int g = 0; g < G; g++) {
        s0[g] = n_sub * g / G;
        s1[g] = n_sub * (g + 1) / G;
        width = std::max(width, s1[g] - s0[g]);
    }
    uint8_t zeroes[33][32];
    inf_merkle_zeroes(m->ctx[0], arity, &zeroes[0][0]);
    const size_t gather = (size_t)width * 32;
    cudaError_t e = cudaSuccess;
    // Every exit after work has been queued goes through here: nothing may still be reading the
    // caller's leaves, or using the contexts' buffers, when the call returns.
    auto finish = [&](int code, const char* what) {
        for (int g = 0; g < G; g++) {
            const int drc = inf_internal_drain(m->ctx[g]);
            if (!code && drc) code = drc;
        }
        if (code == INF_ERR_CUDA && what) m->last_error = std::string(what) + ": " + cudaGetErrorString(e);
        else if (code) m->last_error = inf_strerror(code);
        return code;
    };
    if (gather > m->gather_bytes) {
        // drop the old buffers first and forget them, so that a failed allocation leaves the object
        // with no buffers and gather_bytes == 0 rather than with dangling pointers
        m->gather_bytes = 0;
        for (int g = 0; g < G; g++) {
            cudaSetDevice(m->devices[g]);
            if (m->send[g]) cudaFree(m->send[g]);
            if (m->recv[g]) cudaFree(m->recv[g]);
            m->send[g] = m->recv[g] = nullptr;
        }
        for (int g = 0; g < G; g++) {
            cudaSetDevice(m->devices[g]);
            if ((e = cudaMalloc(&m->send[g], gather)) != cudaSuccess || (e = cudaMalloc(&m->recv[g], gather * G)) != cudaSuccess) {
                for (int h = 0; h <= g; h++) {
                    cudaSetDevice(m->devices[h]);
                    if (m->send[h]) cudaFree(m->send[h]);
                    if (m->recv[h]) cudaFree(m->recv[h]);
                    m->send[h] = m->recv[h] = nullptr;
                }
                cudaGetLastError();
                return finish(e == cudaErrorMemoryAllocation ? INF_ERR_OUT_OF_MEMORY : INF_ERR_CUDA, "cudaMalloc");
            }
        }
        m->gather_bytes = gather;
    }
    std::vector<cudaStream_t> st(G);
    for (int g = 0; g < G; g++) {
        void* sp = nullptr;
        inf_internal_stream(m->ctx[g], &sp);
        st[g] = (cudaStream_t)sp;
    }
    // 1. per device, on a thread of its own: upload the slice in chunks, reduce it to level k into the send buffer
    std::vector<int> rcs(G, INF_OK);
    auto reduce_on = [&](int g) {
        if (s1[g] == s0[g]) return;
        cudaSetDevice(m->devices[g]);
        const uint64_t lo_log = s0[g] * w, hi_log = std::min<uint64_t>(s1[g] * w, n_total);
        const uint64_t lo = lo_log >= shift ? lo_log - shift : 0, hi = hi_log >= shift ? hi_log - shift : 0;
        const uint64_t cnt = hi > lo ? hi - lo : 0;
        const uint64_t sh = s0[g] == 0 ? shift : 0;
        if (cudaMemsetAsync(m->send[g], 0, gather, st[g]) != cudaSuccess) {
            rcs[g] = INF_ERR_CUDA;
            return;
        }
        uint64_t got = 0;
        if (k == 0) {                                            // the run is the leaves themselves
            // (only trees too small to be worth sharding: fewer than 64 x G leaves)
            cudaError_t ce = cudaSuccess;
            for (uint64_t i = 0; i < sh && ce == cudaSuccess; i++)
                ce = cudaMemcpyAsync((char*)m->send[g] + 32 * i, zeroes[0], 32, cudaMemcpyHostToDevice, st[g]);
            if (ce == cudaSuccess && cnt)
                ce = cudaMemcpyAsync((char*)m->send[g] + 32 * sh, leaves + lo * 32, cnt * 32, cudaMemcpyHostToDevice, st[g]);
            rcs[g] = ce == cudaSuccess ? INF_OK : INF_ERR_CUDA;
            got = cnt + sh;
        } else {
            rcs[g] = inf_internal_tree_reduce_host(m->ctx[g], arity, k, sh, leaves + lo * 32, cnt, m->send[g], &got);
        }
        if (!rcs[g] && got != s1[g] - s0[g]) rcs[g] = INF_ERR_MERGE_FAILED;
    };
    {
        std::vector<std::thread> th;
        for (int g = 1; g < G; g++) th.emplace_back(reduce_on, g);
        reduce_on(0);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < G; g++)
        if (rcs[g]) return finish(rcs[g], "device reduction");
    // 2. exchange the subtree roots
    if (G > 1) {
        if (m->use_nccl) {
            g_nccl.GroupStart();
            bool bad = false;
            for (int g = 0; g < G; g++) {
                cudaSetDevice(m->devices[g]);
                if (g_nccl.AllGather(m->send[g], m->recv[g], gather, /*ncclChar*/ 0, m->comms[g], st[g]) != 0) bad = true;
            }
            if (g_nccl.GroupEnd() != 0 || bad) return finish(INF_ERR_NCCL, nullptr);
        } else {
            // peer copies into device 0's receive buffer, ordered after each producer
            for (int g = 0; g < G; g++) {
                cudaSetDevice(m->devices[g]);
                if ((e = cudaStreamSynchronize(st[g])) != cudaSuccess) return finish(INF_ERR_CUDA, "cudaStreamSynchronize");
            }
            cudaSetDevice(m->devices[0]);
            for (int g = 0; g < G; g++)
                if ((e = cudaMemcpyPeerAsync((char*)m->recv[0] + gather * g, m->devices[0], m->send[g], m->devices[g],
                                             gather, st[0])) != cudaSuccess)
                    return finish(INF_ERR_CUDA, "cudaMemcpyPeerAsync");
        }
    }
    // 3. device 0: compact the runs and finish the top levels
    cudaSetDevice(m->devices[0]);
    void* d_nodes = nullptr;
    int rc = inf_internal_grow_io(m->ctx[0], 1, std::max<uint64_t>(n_sub, 1) * 32 + 32, &d_nodes);
    if (rc) return finish(rc, nullptr);
    uint64_t off = 0;
    for (int g = 0; g < G; g++) {
        const uint64_t c = s1[g] - s0[g];
        if (!c) continue;
        const void* src = G > 1 ? (const void*)((char*)m->recv[0] + gather * g) : (const void*)m->send[0];
        if ((e = cudaMemcpyAsync((char*)d_nodes + off * 32, src, c * 32, cudaMemcpyDeviceToDevice, st[0])) != cudaSuccess)
            return finish(INF_ERR_CUDA, "cudaMemcpyAsync D2D");
        off += c;
    }
    void* d_root = (char*)d_nodes + n_sub * 32;
    uint64_t got = 0;
    rc = inf_tree_reduce_dev(m->ctx[0], arity, k, rdepth - k, 0, d_nodes, n_sub, d_root, &got, st[0]);
    if (rc) return finish(rc, nullptr);
    if (got != 1) return finish(INF_ERR_MERGE_FAILED, nullptr);
    uint8_t r[32];
    if ((e = cudaMemcpyAsync(r, d_root, 32, cudaMemcpyDeviceToHost, st[0])) != cudaSuccess) return finish(INF_ERR_CUDA, "D2H root");
    // the root is there when device 0's stream is; with NCCL every device took part in the
    // collective on its own stream: drain them all
    if ((rc = finish(INF_OK, nullptr))) return rc;
    if (root) memcpy(root, r, 32);
    if (has_root) *has_root = 1;
    return completed ? INF_ERR_TREE_ALREADY_MERGED : INF_OK;
}

// n independent hashes split into contiguous slices, one per device, run
// concurrently (host-buffer pipelines of each context on their own streams).
int inf_multi_poseidon_hash_batch(inf_multi* m, uint32_t n_inputs, uint32_t flags, const uint8_t* domain_tag,
                                  const uint8_t* in, uint64_t n, uint8_t* out) {
    if (!m) return INF_ERR_NULL_POINTER;
    const int G = (int)m->devices.size();
    // The single-context call is synchronous, so slices run back to back from one
    // host thread unless the caller threads them; for the common large-batch case
    // use one std::thread per device.
    std::vector<int> rcs(G, 0);
    std::vector<std::pair<uint64_t, uint64_t>> rng(G);
    for (int g = 0; g < G; g++) rng[g] = {n * g / G, n * (g + 1) / G};
    struct Job { inf_multi* m; int g; uint32_t k, flags; const uint8_t *tag, *in; uint8_t* out; uint64_t lo, hi; int* rc; };
    std::vector<Job> jobs;
    for (int g = 0; g < G; g++) jobs.push_back({m, g, n_inputs, flags, domain_tag, in, out, rng[g].first, rng[g].second, &rcs[g]});
    auto run = [](void* p) -> void* {
        Job* j = (Job*)p;
        cudaSetDevice(j->m->devices[j->g]);
        *j->rc = inf_poseidon_hash_batch(j->m->ctx[j->g], j->k, j->flags, j->tag, j->in + j->lo * j->k * 32,
                                         j->hi - j->lo, j->out + j->lo * 32);
        return nullptr;
    };
    DeviceRestore restore;
    std::vector<pthread_t> th(G);
    std::vector<bool> started(G, false);
    for (int g = 1; g < G; g++) started[g] = pthread_create(&th[g], nullptr, run, &jobs[g]) == 0;
    run(&jobs[0]);
    for (int g = 1; g < G; g++) {
        if (started[g]) pthread_join(th[g], nullptr);
        else run(&jobs[g]);                      // no thread to be had: do the slice here
    }
    for (int g = 0; g < G; g++)
        if (rcs[g]) return rcs[g];
    return INF_OK;
}

}  // extern "C"
