// mc_tessellate_multi: z-slab sharding of one meshing job over several devices driven by ONE
// process (the call a single-process host such as the reference's `main`,
// src/main.rs:74-82, makes).  One host thread per device inside the call; the only exchanges are
//   * the per-slab (vertex, tet) counts, in host memory, and
//   * slab i's id table of the shared lattice plane + the coordinates of its top tile layer,
//     device to device (cudaMemcpyPeerAsync, NVLink when the devices are peers), ordered by an event.
// The same phases, in the same order, as the multi-process driver (moosecad_b200/slabs.py).
#include <condition_variable>
#include <mutex>
#include <thread>

#include "mc_internal.hpp"

using namespace mc;

namespace {

class Barrier {  // C++17: no std::barrier
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0, phase = 0;
  public:
    explicit Barrier(int count) : n(count) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int ph = phase;
        if (++waiting == n) { waiting = 0; ++phase; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != ph; });
    }
};

// contiguous z ranges with (nearly) equal summed cost, at least one layer each (slabs.py partition_balanced)
std::vector<std::pair<uint64_t, uint64_t>> partition_balanced(const std::vector<double>& cost, int world) {
    const uint64_t n = cost.size();
    std::vector<std::pair<uint64_t, uint64_t>> out;
    double total = 0.0;
    for (double c : cost) total += c;
    if (world <= 1 || n <= (uint64_t)world || !(total > 0.0)) {
        const uint64_t base = n / (uint64_t)world, rem = n % (uint64_t)world;
        uint64_t z = 0;
        for (int r = 0; r < world; ++r) {
            const uint64_t k = base + ((uint64_t)r < rem ? 1 : 0);
            out.emplace_back(z, z + k);
            z += k;
        }
        return out;
    }
    uint64_t z = 0;
    double acc = 0.0;
    for (int r = 0; r < world; ++r) {
        const int remaining = world - r;
        if (remaining == 1) { out.emplace_back(z, n); break; }
        const double target = (total - acc) / remaining;
        uint64_t z1 = z;
        double s = 0.0;
        while (z1 < n - (uint64_t)(remaining - 1) && (z1 == z || s + 0.5 * cost[z1] <= target)) { s += cost[z1]; ++z1; }
        out.emplace_back(z, z1);
        acc += s;
        z = z1;
    }
    return out;
}

}  // namespace

extern "C" int mc_tessellate_multi(mc_ctx* const* ctxs, int n, const mc_tape* volume, int32_t root, double resolution,
                                   double relative_error, const mc_options* opts, mc_mesh** out) {
    if (!ctxs || n <= 0 || !volume || !out) return fail(MC_ERR_ARG, "mc_tessellate_multi: bad argument");
    for (int i = 0; i < n; ++i) {
        out[i] = nullptr;
        if (!ctxs[i]) return fail(MC_ERR_ARG, "mc_tessellate_multi: null context");
        for (int j = 0; j < i; ++j)
            if (ctxs[j] == ctxs[i]) return fail(MC_ERR_ARG, "mc_tessellate_multi: a context may drive one slab only");
    }
    if (n == 1) return mc_tessellate(ctxs[0], volume, root, resolution, relative_error, opts, &out[0]);

    // slab boundaries from the load estimate (deterministic; one interval pass on the first device)
    double bbox[6];
    int rc = mc_tape_bbox(volume, root, bbox);
    if (rc != MC_OK) return rc;
    const double dz = std::ceil((bbox[5] - bbox[2]) / resolution);
    if (!(resolution > 0.0) || !(dz >= 0.0) || !(dz < 4294967295.0))
        return fail(MC_ERR_BBOX, "mc_tessellate_multi: the volume's bounding box is not finite");
    const uint64_t nz = (uint64_t)dz;
    if (nz < (uint64_t)n) return fail(MC_ERR_ARG, "mc_tessellate_multi: fewer cell layers than devices");
    std::vector<double> cost(nz, 1.0);
    rc = mc_estimate_layer_costs(ctxs[0], volume, root, resolution, cost.data(), nz);
    if (rc != MC_OK) return rc;
    const auto slabs = partition_balanced(cost, n);

    struct Rank {
        int rc = MC_OK;
        std::string err;
        uint64_t counts[6] = {0, 0, 0, 0, 0, 0};
        uint64_t top_first = 0, top_count = 0;
        cudaEvent_t ready = nullptr;  // plane table + top-layer coordinates of this slab are complete
    };
    std::vector<Rank> rk((size_t)n);
    Barrier bar(n);
    auto failed_any = [&]() { for (auto& r : rk) if (r.rc != MC_OK && r.rc != MC_ERR_DIVERGED) return true; return false; };

    auto worker = [&](int i) {
        Rank& me = rk[(size_t)i];
        mc_ctx* c = ctxs[i];
        auto step = [&](int code) { if (code != MC_OK && me.rc == MC_OK) { me.rc = code; me.err = mc_last_error(); } return code == MC_OK; };
        mc_slab slab{slabs[(size_t)i].first, slabs[(size_t)i].second};
        mc_mesh* m = nullptr;
        if (step(mc_tessellate_begin(c, volume, root, resolution, relative_error, &slab, opts, &m))) {
            out[i] = m;
            step(mc_mesh_counts(m, me.counts));
            step(mc_mesh_upper_layer_vertices(m, &me.top_first, &me.top_count));
            step(mc_tessellate_emit_vertices_local(m));  // base-independent half, overlaps the wait below
        }
        bar.arrive_and_wait();  // every slab's counts are known
        const bool go = !failed_any();
        uint64_t vbase = 0, tbase = 0;
        for (int r = 0; r < i; ++r) { vbase += rk[(size_t)r].counts[0]; tbase += rk[(size_t)r].counts[1]; }
        if (go) {
            if (step(mc_tessellate_emit_vertices(m, vbase))) {
                cudaError_t e = cudaSetDevice(c->device);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&me.ready, cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventRecord(me.ready, c->stream);
                if (e != cudaSuccess) step(cuda_fail(e, "mc_tessellate_multi: event"));
            }
        }
        bar.arrive_and_wait();  // every slab's table is queued and its event recorded
        if (go && !failed_any()) {
            const void* lower = nullptr;
            if (i > 0) {
                // pull the table and the coordinates from the slab below onto this device
                mc_mesh* below = out[i - 1];
                const Rank& rb = rk[(size_t)i - 1];
                void* src_table = nullptr;
                uint64_t n_entries = 0;
                void* src_xyz = nullptr;
                bool ok = step(mc_mesh_upper_plane_table(below, &src_table, &n_entries));
                if (ok && rb.top_count) ok = step(mc_mesh_upper_layer_coords(below, &src_xyz));
                if (ok) ok = step(c->alloc(std::max<uint64_t>(n_entries, 1) * 8, m->own_lower_table));
                if (ok) ok = step(c->alloc(std::max<uint64_t>(rb.top_count, 1) * 24, m->own_lower_coords));
                if (ok) {
                    cudaError_t e = cudaSetDevice(c->device);
                    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, rb.ready, 0);
                    if (e == cudaSuccess && n_entries)
                        e = cudaMemcpyPeerAsync(m->own_lower_table.p, c->device, src_table, below->ctx->device, n_entries * 8, c->stream);
                    if (e == cudaSuccess && rb.top_count)
                        e = cudaMemcpyPeerAsync(m->own_lower_coords.p, c->device, src_xyz, below->ctx->device, rb.top_count * 24, c->stream);
                    if (e != cudaSuccess) ok = step(cuda_fail(e, "mc_tessellate_multi: peer copy"));
                }
                if (ok) {
                    uint64_t vb_below = 0;
                    for (int r = 0; r < i - 1; ++r) vb_below += rk[(size_t)r].counts[0];
                    ok = step(mc_tessellate_set_lower_coords(m, rb.top_count ? m->own_lower_coords.p : nullptr,
                                                             vb_below + rb.top_first, rb.top_count));
                    lower = m->own_lower_table.p;
                }
            }
            if (me.rc == MC_OK) {
                const int code = mc_tessellate_emit_tets(m, tbase, lower);  // ends with a stream synchronisation
                if (code != MC_OK) { me.rc = code; me.err = mc_last_error(); }
            }
        }
        bar.arrive_and_wait();  // nobody's source buffers are needed any more
        if (me.ready) { cudaSetDevice(c->device); cudaEventDestroy(me.ready); me.ready = nullptr; }
    };

    std::vector<std::thread> threads;
    for (int i = 1; i < n; ++i) threads.emplace_back(worker, i);
    worker(0);
    for (auto& t : threads) t.join();

    int worst = MC_OK;
    std::string msg;
    for (auto& r : rk)
        if (r.rc != MC_OK && (worst == MC_OK || worst == MC_ERR_DIVERGED)) { worst = r.rc; msg = r.err; }
    if (worst != MC_OK && worst != MC_ERR_DIVERGED) {
        for (int i = 0; i < n; ++i) { mc_mesh_free(out[i]); out[i] = nullptr; }
        return fail(worst, "mc_tessellate_multi: " + msg);
    }
    if (worst == MC_ERR_DIVERGED) return fail(worst, msg);
    return MC_OK;
}
