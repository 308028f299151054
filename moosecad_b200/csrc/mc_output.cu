// Views of an emitted mesh that are not the mesh crate's own layout:
//   * mc_mesh_checksum: order- and numbering-independent 64-bit checksums of the vertices and
//     tets, computed on the device (parity of two runs / of N slabs against one without moving
//     the mesh to the host; tests/canon.py holds the numpy mirror used against the oracle).
//   * mc_mesh_exodus_*: the arrays exodus::write builds (transposed coordinates, i64 1-based
//     connectivity, i32 1-based side-set arrays), laid out on the device and streamed to the host.
#include <algorithm>
#include <string>

#include "mc_internal.hpp"
#include "mc_hash.cuh"

using namespace mc;

namespace mc {

static __global__ void __launch_bounds__(256)
k_checksum_vertices(const double* __restrict__ coords, uint64_t nv, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (uint64_t)gridDim.x * blockDim.x)
        acc += vertex_hash(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31u) == 0u && acc) atomicAdd(&out[0], acc);
}

// Vertices owned by the rank below (ids under vertex_base) are looked up in the slice of its
// coordinates that came with the plane table, like k_tet_quality_list does.
template <typename IndexT>
__global__ void __launch_bounds__(256)
k_checksum_tets(const double* __restrict__ coords, const IndexT* __restrict__ tets, uint64_t nt, long long vertex_base,
                const double* __restrict__ lower_coords, long long lower_first, long long lower_count,
                unsigned long long* __restrict__ out) {
    unsigned long long acc = 0ull, missing = 0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h[4];
        bool have = true;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const long long gid = (long long)tets[4 * i + m];
            const double* src = coords + 3 * (gid - vertex_base);
            if (gid < vertex_base) {
                const long long lid = gid - lower_first;
                if (lower_coords == nullptr || lid < 0 || lid >= lower_count) { have = false; src = coords; }
                else src = lower_coords + 3 * lid;
            }
            h[m] = vertex_hash(src[0], src[1], src[2]);
        }
        if (have) acc += tet_hash(h[0], h[1], h[2], h[3]);
        else ++missing;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        missing += __shfl_xor_sync(0xffffffffu, missing, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (acc) atomicAdd(&out[1], acc);
        if (missing) atomicAdd(&out[2], missing);
    }
}

}  // namespace mc

extern "C" int mc_mesh_checksum(mc_mesh* m, uint64_t out[4]) {
    if (!m || !out) return fail(MC_ERR_ARG, "mc_mesh_checksum: bad argument");
    if (m->phase < 3) return fail(MC_ERR_STATE, "mc_mesh_checksum: mesh not emitted");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    DevBuf acc;
    int rc = c->alloc(4 * 8, acc);
    if (rc != MC_OK) return rc;
    cudaError_t e = cudaMemsetAsync(acc.p, 0, 32, s);
    const unsigned grid = (unsigned)c->sm_count * 8u;
    if (e == cudaSuccess && m->n_verts) {
        k_checksum_vertices<<<grid, 256, 0, s>>>((const double*)m->d_coords.p, m->n_verts, (unsigned long long*)acc.p);
        c->launches++;
    }
    if (e == cudaSuccess && m->n_tets) {
        if (m->index_bytes == 4)
            k_checksum_tets<uint32_t><<<grid, 256, 0, s>>>((const double*)m->d_coords.p, (const uint32_t*)m->d_tets.p, m->n_tets,
                                                           (long long)m->vertex_base, m->lower_coords, (long long)m->lower_first,
                                                           (long long)m->lower_count, (unsigned long long*)acc.p);
        else
            k_checksum_tets<unsigned long long><<<grid, 256, 0, s>>>(
                (const double*)m->d_coords.p, (const unsigned long long*)m->d_tets.p, m->n_tets, (long long)m->vertex_base,
                m->lower_coords, (long long)m->lower_first, (long long)m->lower_count, (unsigned long long*)acc.p);
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, acc.p, 32, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    c->release(acc);
    if (e != cudaSuccess) return cuda_fail(e, "mc_mesh_checksum");
    return MC_OK;
}

// ---------------------------------------------------------------------------------------------
// exodus::write's arrays (exodus/src/lib.rs:55-63 coord, :84-96 connect1, :118-132 elem_ss / side_ss),
// produced on the device in the writer's layout and streamed into the caller's HOST buffers, so that the
// consumer does straight `put_values` calls instead of looping add_vertex / add_elem over 10^9 items.
// Chunks go through two device staging buffers: the kernel that lays out chunk k + 1 runs while
// chunk k crosses the bus on the copy stream.
namespace mc {

static __global__ void __launch_bounds__(256)
k_exodus_coord(const double* __restrict__ xyz, uint64_t first, uint64_t n, double* __restrict__ stage /* [3][n] */) {
    // coalesced on both sides: thread e handles double e of the AoS slice
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 3 * n; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = e / 3, k = e - 3 * v;
        stage[k * n + v] = xyz[3 * first + e];
    }
}

template <typename IndexT>
__global__ void __launch_bounds__(256)
k_exodus_connect(const IndexT* __restrict__ tets, uint64_t first, uint64_t n, long long* __restrict__ stage) {
    // e.node(j) as i64 + 1
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 4 * n; e += (uint64_t)gridDim.x * blockDim.x)
        stage[e] = (long long)tets[4 * first + e] + 1;
}

static __global__ void __launch_bounds__(256)
k_exodus_sideset(const unsigned long long* __restrict__ pairs, uint64_t n, int* __restrict__ elem_ss, int* __restrict__ side_ss,
                 unsigned long long* __restrict__ overflow) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = pairs[2 * i] + 1ull, s = pairs[2 * i + 1] + 1ull;
    if (e > 2147483647ull) atomicAdd(overflow, 1ull);  // `*elem as i32 + 1` would wrap in the reference
    elem_ss[i] = (int)e;
    side_ss[i] = (int)s;
}

// double-buffered pipeline: lay out `count` items in chunks with `launch(first, n, stage)`, copy every
// chunk to the host with `copy(first, n, stage)` on the copy stream
template <typename Launch, typename Copy>
static int stream_out(mc_ctx* c, uint64_t count, size_t item_bytes, Launch launch, Copy copy) {
    if (count == 0) return MC_OK;
    const uint64_t chunk = std::max<uint64_t>(1, (size_t)(96u << 20) / item_bytes);
    DevBuf stage[2];
    cudaEvent_t filled[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    int rc = MC_OK;
    cudaError_t e = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && rc == MC_OK && e == cudaSuccess; ++b) {
        rc = c->alloc(std::min<uint64_t>(chunk, count) * item_bytes, stage[b]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&filled[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&drained[b], cudaEventDisableTiming);
    }
    int k = 0;
    for (uint64_t first = 0; first < count && rc == MC_OK && e == cudaSuccess; first += chunk, ++k) {
        const int b = k & 1;
        const uint64_t n = std::min<uint64_t>(chunk, count - first);
        if (k >= 2) e = cudaStreamWaitEvent(c->stream, drained[b], 0);  // the copy of chunk k - 2 has left this buffer
        if (e == cudaSuccess) { launch(first, n, stage[b].p); c->launches++; e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaEventRecord(filled[b], c->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(copy_stream, filled[b], 0);
        if (e == cudaSuccess) e = copy(first, n, stage[b].p, copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(drained[b], copy_stream);
    }
    if (copy_stream) { cudaError_t e2 = cudaStreamSynchronize(copy_stream); if (e == cudaSuccess) e = e2; }
    { cudaError_t e2 = cudaStreamSynchronize(c->stream); if (e == cudaSuccess) e = e2; }
    for (int b = 0; b < 2; ++b) {
        c->release(stage[b]);
        if (filled[b]) cudaEventDestroy(filled[b]);
        if (drained[b]) cudaEventDestroy(drained[b]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (rc != MC_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "exodus layout");
    return MC_OK;
}

}  // namespace mc

extern "C" int mc_mesh_exodus_coord(mc_mesh* m, double* coord, uint64_t num_nodes_total) {
    if (!m || !coord) return fail(MC_ERR_ARG, "mc_mesh_exodus_coord: bad argument");
    if (m->phase < 3) return fail(MC_ERR_STATE, "mc_mesh_exodus_coord: mesh not emitted");
    if (m->vertex_base + m->n_verts > num_nodes_total) return fail(MC_ERR_ARG, "mc_mesh_exodus_coord: num_nodes_total is too small");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    const double* xyz = (const double*)m->d_coords.p;
    const unsigned grid = (unsigned)c->sm_count * 8u;
    return stream_out(
        c, m->n_verts, 24,
        [&](uint64_t first, uint64_t n, void* stage) { k_exodus_coord<<<grid, 256, 0, c->stream>>>(xyz, first, n, (double*)stage); },
        [&](uint64_t first, uint64_t n, void* stage, cudaStream_t cs) {
            // row k of coord[num_dim][num_nodes], columns of this slab's vertices
            cudaError_t e = cudaSuccess;
            for (int k = 0; k < 3 && e == cudaSuccess; ++k)
                e = cudaMemcpyAsync(coord + (size_t)k * num_nodes_total + m->vertex_base + first, (const double*)stage + (size_t)k * n,
                                    n * 8, cudaMemcpyDeviceToHost, cs);
            return e;
        });
}

extern "C" int mc_mesh_exodus_connect(mc_mesh* m, int64_t* connect1) {
    if (!m || !connect1) return fail(MC_ERR_ARG, "mc_mesh_exodus_connect: bad argument");
    if (m->phase < 3) return fail(MC_ERR_STATE, "mc_mesh_exodus_connect: mesh not emitted");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    const unsigned grid = (unsigned)c->sm_count * 8u;
    const void* tets = m->d_tets.p;
    const int ib = m->index_bytes;
    return stream_out(
        c, m->n_tets, 32,
        [&](uint64_t first, uint64_t n, void* stage) {
            if (ib == 4) k_exodus_connect<uint32_t><<<grid, 256, 0, c->stream>>>((const uint32_t*)tets, first, n, (long long*)stage);
            else k_exodus_connect<unsigned long long><<<grid, 256, 0, c->stream>>>((const unsigned long long*)tets, first, n, (long long*)stage);
        },
        [&](uint64_t first, uint64_t n, void* stage, cudaStream_t cs) {
            return cudaMemcpyAsync(connect1 + 4 * first, stage, n * 32, cudaMemcpyDeviceToHost, cs);
        });
}

extern "C" int mc_mesh_exodus_sideset(mc_mesh* m, int index, int32_t* elem_ss, int32_t* side_ss) {
    if (!m || index < 0 || (size_t)index >= m->sidesets.size()) return fail(MC_ERR_ARG, "mc_mesh_exodus_sideset: bad index");
    const SideSetDev& ss = m->sidesets[(size_t)index];
    if (ss.n == 0) return MC_OK;
    if (!elem_ss || !side_ss) return fail(MC_ERR_ARG, "mc_mesh_exodus_sideset: null buffer");
    mc_ctx* c = m->ctx;
    MC_CUDA(cudaSetDevice(c->device));
    DevBuf de, ds, dov;
    int rc = c->alloc(ss.n * 4, de);
    if (rc == MC_OK) rc = c->alloc(ss.n * 4, ds);
    if (rc == MC_OK) rc = c->alloc(8, dov);
    unsigned long long overflow = 0ull;
    cudaError_t e = cudaSuccess;
    if (rc == MC_OK) {
        e = cudaMemsetAsync(dov.p, 0, 8, c->stream);
        k_exodus_sideset<<<(unsigned)((ss.n + 255) / 256), 256, 0, c->stream>>>((const unsigned long long*)ss.d_pairs.p, ss.n, (int*)de.p,
                                                                             (int*)ds.p, (unsigned long long*)dov.p);
        c->launches++;
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(elem_ss, de.p, ss.n * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(side_ss, ds.p, ss.n * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&overflow, dov.p, 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    c->release(de); c->release(ds); c->release(dov);
    if (rc != MC_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "mc_mesh_exodus_sideset");
    if (overflow) return fail(MC_ERR_LIMIT, "mc_mesh_exodus_sideset: " + std::to_string(overflow) +
                                            " element numbers exceed the i32 range of elem_ss (the reference's `as i32` would wrap)");
    return MC_OK;
}
