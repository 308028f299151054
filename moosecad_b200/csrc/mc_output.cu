// Views of an emitted mesh that are not the mesh crate's own layout:
//   * mc_mesh_checksum: order- and numbering-independent 64-bit checksums of the vertices and
//     tets, computed on the device (parity of two runs / of N slabs against one without moving
//     the mesh to the host; tests/canon.py holds the numpy mirror used against the oracle).
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
