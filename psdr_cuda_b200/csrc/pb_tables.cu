// psdr-b200: the tables of Scene::configure that depend on vertex positions, built on the device — primary-edge list of a sensor
// (src/sensor/perspective.cpp:39-111), secondary-edge table (src/shape/mesh.cpp:251-264, src/scene/scene.cpp:219-235), face-area
// and environment-map distributions (mesh.cpp:238-249, src/emitter/envmap.cpp:10-26), scene bounds (scene.cpp:88-119).
//
// In an optimisation loop the reference rebuilds all of them on the GPU through Enoki at every configure(); round 1 of this repo
// did it on the host (download the vertices, loop over 10^5 edges, upload 8 MB of records: 10 ms per configure, 0.1-0.4 s with an
// environment map). Here every table is a few kernels over the resident triangle table / world-space vertices:
//   flags (one thread per mesh edge)  ->  exclusive scan of the flags (order-preserving compaction: the reference keeps mesh order,
//   and the order decides which edge a sample picks)  ->  records + lengths  ->  cmf.
// The cmfs are *sequential* fp32 running sums like the reference's / the oracle's, so that discrete sampling picks the same entry for
// the same sample: one thread adds, its block stages the operands through shared memory (4 dependent-add cycles per entry: 0.2 ms for
// 10^5 edges, 4 ms for the 2 M cells of a 1024 x 512 environment map).
#include "pb_kernels.h"
#include "pb_shade.cuh"

namespace pb {

static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

// ---- sequential inclusive prefix sum (fp32, left to right), total to sum_out -------------------------------------------------
__global__ void __launch_bounds__(1024) k_seq_cmf(long long n, const float *__restrict__ pmf, float *__restrict__ cmf, float *__restrict__ sum_out, const int *__restrict__ n_dev) {
    __shared__ __align__(16) float s_buf[2][4096];
    if (n_dev) n = *n_dev;
    float acc = 0.f;
    const int tid = threadIdx.x;
    const long long chunks = (n + 4095) / 4096;
    for (long long c = 0; c < chunks; ++c) {
        float *buf = s_buf[c & 1];
        const long long base = c * 4096;
        for (int k = tid; k < 4096; k += 1024) buf[k] = (base + k < n) ? pmf[base + k] : 0.f;
        __syncthreads();
        if (tid == 0) {   // the one dependent chain; operands come in register batches so that only the adds wait for each other
            const int m = (int)min((long long)4096, n - base);
            int k = 0;
            float4 *b4 = reinterpret_cast<float4 *>(buf);
            for (; k + 16 <= m; k += 16) {   // 128-bit shared-memory accesses: 4 loads + 16 dependent adds + 4 stores per 16 entries
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = b4[(k >> 2) + j];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc = __fadd_rn(acc, v[j].x); v[j].x = acc; acc = __fadd_rn(acc, v[j].y); v[j].y = acc;
                    acc = __fadd_rn(acc, v[j].z); v[j].z = acc; acc = __fadd_rn(acc, v[j].w); v[j].w = acc;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) b4[(k >> 2) + j] = v[j];
            }
            for (; k < m; ++k) { acc = __fadd_rn(acc, buf[k]); buf[k] = acc; }
        }
        __syncthreads();
        for (int k = tid; k < 4096; k += 1024) if (base + k < n) cmf[base + k] = buf[k];
    }
    if (tid == 0 && sum_out) *sum_out = acc;
}
void launch_seq_cmf(cudaStream_t st, long long n, const float *pmf, float *cmf, float *sum_out, const int *n_dev) {
    k_seq_cmf<<<1, 1024, 0, st>>>(n, pmf, cmf, sum_out, n_dev);
}

// ---- order-preserving compaction: exclusive scan of 0/1 flags -------------------------------------------------------------------
constexpr int kScanTile = 4096;   // flags per block
__global__ void __launch_bounds__(1024) k_scan_tiles(int n, const unsigned char *__restrict__ flags, int *__restrict__ local, int *__restrict__ tile_sum) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, base = blockIdx.x * kScanTile + tid * 4;
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? flags[base + k] : 0; sum += v[k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += x; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = s_warp[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += x; }
        s_warp[tid] = w;
    }
    __syncthreads();
    int excl = incl - sum + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < n) local[base + k] = excl; excl += v[k]; }
    if (tid == 1023) tile_sum[blockIdx.x] = s_warp[31];
}
__global__ void __launch_bounds__(1024) k_scan_tile_sums(int tiles, int *__restrict__ tile_sum, int *__restrict__ total) {   // in place, exclusive; tiles <= 1024 * 64
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    const int per = (tiles + 1023) / 1024;
    int sum = 0;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < tiles) sum += tile_sum[i]; }
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { const int x = tid >= o ? s_part[tid - o] : 0; __syncthreads(); s_part[tid] += x; __syncthreads(); }
    int base = s_part[tid] - sum;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < tiles) { const int t = tile_sum[i]; tile_sum[i] = base; base += t; } }
    if (tid == 1023) *total = s_part[1023];
}

// edge source record (topology, uploaded once): v0 v1 f0 f1(-1: boundary) v2(opposite vertex) mesh
struct EdgeSrc { int v0, v1, f0, f1, v2, mesh; };

PB_D float3 tri_p0(const TriRec *tri, int t) { return f3(ldg4(reinterpret_cast<const float4 *>(tri + t))); }
PB_D float3 tri_fn(const TriRec *tri, int t) { return f3(ldg4(reinterpret_cast<const float4 *>(tri + t) + 6)); }
PB_D float3 vert3(const float *v, int i) { return f3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

// perspective.cpp:47-66: which edges can be silhouettes from this camera
__global__ void __launch_bounds__(256) k_prim_edge_flags(int n, const EdgeSrc *__restrict__ es, const TriRec *__restrict__ tri, const MeshRec *__restrict__ meshes,
                                                         float3 cam, unsigned char *__restrict__ flags, int *__restrict__ mesh_kept) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EdgeSrc e = es[i];
    const MeshRec m = meshes[e.mesh];
    const bool interior = e.f1 >= 0;
    const float3 e0 = normalize(cam - tri_p0(tri, m.face_offset + e.f0)), n0 = tri_fn(tri, m.face_offset + e.f0);
    float3 e1 = f3(0.f), n1 = f3(0.f);
    if (interior) { e1 = normalize(cam - tri_p0(tri, m.face_offset + e.f1)); n1 = tri_fn(tri, m.face_offset + e.f1); }
    bool keep;
    if (m.flags & 1) keep = !(interior && ((dot(e0, n0) < kEpsilon && dot(e1, n1) < kEpsilon) || dot(n0, n1) > 1.f - kEpsilon));
    else keep = !interior || ((dot(e0, n0) > kEpsilon) != (dot(e1, n1) > kEpsilon));
    flags[i] = keep ? 1 : 0;
    if (keep) atomicAdd(mesh_kept + e.mesh, 1);
}
// perspective.cpp:68-111: project the kept edges, unit normal and length in film space
__global__ void __launch_bounds__(256) k_prim_edge_write(int n, const EdgeSrc *__restrict__ es, const unsigned char *__restrict__ flags, const int *__restrict__ local,
                                                         const int *__restrict__ tile_off, const float *const *__restrict__ vworld, Mat4 w2s,
                                                         PrimEdgeRec *__restrict__ recs, float *__restrict__ pmf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const int pos = local[i] + tile_off[i / kScanTile];
    const EdgeSrc e = es[i];
    const float *v = vworld[e.mesh];
    const float3 q0 = transform_pos(w2s, vert3(v, e.v0)), q1 = transform_pos(w2s, vert3(v, e.v1));
    PrimEdgeRec r;
    r.p0x = q0.x; r.p0y = q0.y; r.p1x = q1.x; r.p1y = q1.y;
    const float ex = sub_rn(q1.x, q0.x), ey = sub_rn(q1.y, q0.y);
    const float len = sqrt_rn(fma_rn(ex, ex, mul_rn(ey, ey)));
    r.nx = -div_rn(ey, len); r.ny = div_rn(ex, len); r.len = len; r.pad = 0.f;
    r.mesh = e.mesh; r.v0 = e.v0; r.v1 = e.v1; r.pad2 = 0;
    recs[pos] = r;
    pmf[pos] = len;
}

// mesh.cpp:251-264: every edge that is not flat (boundary edges always qualify: n1 = 0)
__global__ void __launch_bounds__(256) k_sec_edge_flags(int n, const EdgeSrc *__restrict__ es, const TriRec *__restrict__ tri, const MeshRec *__restrict__ meshes,
                                                        unsigned char *__restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EdgeSrc e = es[i];
    const int fo = meshes[e.mesh].face_offset;
    const float3 n0 = tri_fn(tri, fo + e.f0), n1 = e.f1 < 0 ? f3(0.f) : tri_fn(tri, fo + e.f1);
    flags[i] = (dot(n0, n1) < 1.f - kEdgeEpsilon) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_sec_edge_write(int n, const EdgeSrc *__restrict__ es, const unsigned char *__restrict__ flags, const int *__restrict__ local,
                                                        const int *__restrict__ tile_off, const float *const *__restrict__ vworld, const TriRec *__restrict__ tri,
                                                        const MeshRec *__restrict__ meshes, SecEdgeRec *__restrict__ recs, float *__restrict__ pmf, int importance) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const int pos = local[i] + tile_off[i / kScanTile];
    const EdgeSrc e = es[i];
    const float *v = vworld[e.mesh];
    const int fo = meshes[e.mesh].face_offset;
    const bool boundary = e.f1 < 0;
    const float3 p0 = vert3(v, e.v0), q1 = vert3(v, e.v1);
    const float3 e1 = f3(sub_rn(q1.x, p0.x), sub_rn(q1.y, p0.y), sub_rn(q1.z, p0.z));
    const float3 n0 = tri_fn(tri, fo + e.f0), n1 = boundary ? f3(0.f) : tri_fn(tri, fo + e.f1), p2 = vert3(v, e.v2);
    SecEdgeRec r;
    r.a = make_float4(p0.x, p0.y, p0.z, boundary ? 1.f : 0.f);
    r.b = make_float4(e1.x, e1.y, e1.z, __int_as_float(e.mesh));
    r.c = make_float4(n0.x, n0.y, n0.z, __int_as_float(e.v0));
    r.d = make_float4(n1.x, n1.y, n1.z, __int_as_float(e.v1));
    r.e = make_float4(p2.x, p2.y, p2.z, 0.f);
    recs[pos] = r;
    float w = norm(e1);
    if (importance == 1) {   // scene.cpp:230-233 (disabled there): length x exterior dihedral angle, pi for boundary edges
        const float cs = fminf(fmaxf(dot(n0, n1), -1.f), 1.f);
        w *= boundary ? kPi : acosf(cs);
    }
    pmf[pos] = w;
}

static void scan_flags(cudaStream_t st, int n, const unsigned char *flags, int *local, int *tile_sum, int *total) {
    const int tiles = (n + kScanTile - 1) / kScanTile;
    k_scan_tiles<<<tiles, 1024, 0, st>>>(n, flags, local, tile_sum);
    k_scan_tile_sums<<<1, 1024, 0, st>>>(tiles, tile_sum, total);
}

void launch_primary_edge_table(cudaStream_t st, int n, const void *edge_src, const SceneView &S, const float *const *vworld, float3 cam, const Mat4 &w2s,
                               unsigned char *flags, int *local, int *tile_sum, int *mesh_kept, int num_meshes, PrimEdgeRec *recs, float *pmf, float *cmf,
                               int *count_out, float *sum_out) {
    const EdgeSrc *es = static_cast<const EdgeSrc *>(edge_src);
    cudaMemsetAsync(mesh_kept, 0, sizeof(int) * (size_t)num_meshes, st);
    k_prim_edge_flags<<<nblk(n, 256), 256, 0, st>>>(n, es, S.tri, S.meshes, cam, flags, mesh_kept);
    scan_flags(st, n, flags, local, tile_sum, count_out);
    k_prim_edge_write<<<nblk(n, 256), 256, 0, st>>>(n, es, flags, local, tile_sum, vworld, w2s, recs, pmf);
    launch_seq_cmf(st, n, pmf, cmf, sum_out, count_out);
}
void launch_secondary_edge_table(cudaStream_t st, int n, const void *edge_src, const SceneView &S, const float *const *vworld, unsigned char *flags, int *local,
                                 int *tile_sum, SecEdgeRec *recs, float *pmf, float *cmf, int *count_out, float *sum_out, int importance) {
    const EdgeSrc *es = static_cast<const EdgeSrc *>(edge_src);
    k_sec_edge_flags<<<nblk(n, 256), 256, 0, st>>>(n, es, S.tri, S.meshes, flags);
    scan_flags(st, n, flags, local, tile_sum, count_out);
    k_sec_edge_write<<<nblk(n, 256), 256, 0, st>>>(n, es, flags, local, tile_sum, vworld, S.tri, S.meshes, recs, pmf, importance);
    launch_seq_cmf(st, n, pmf, cmf, sum_out, count_out);
}

// ---- environment-map cell masses (envmap.cpp:10-26): luminance of the bilinear lookup at the cell centre x sin(theta) ----------
// sin_theta[j] comes from the host (libm, as the reference / the oracle evaluate it); Bitmap<3>::eval as in bitmap.cpp:43-89
__global__ void __launch_bounds__(256) k_envmap_pmf(int rx, int ry, int w, int h, const float *__restrict__ texel, const float *__restrict__ sin_theta, float *__restrict__ pmf) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)rx * ry) return;
    const int i = (int)(k / ry), j = (int)(k - (long long)i * ry);
    const float ux = div_rn(1.f, (float)rx), uy = div_rn(1.f, (float)ry);
    float u = mul_rn(add_rn((float)i, .5f), ux), v = mul_rn(add_rn((float)j, .5f), uy);
    u = sub_rn(u, floorf(u)); v = sub_rn(v, floorf(v));
    u = mul_rn(u, (float)(w - 1)); v = mul_rn(v, (float)(h - 1));
    int px = (int)floorf(u), py = (int)floorf(v);
    const float w1x = sub_rn(u, (float)px), w1y = sub_rn(v, (float)py), w0x = sub_rn(1.f, w1x), w0y = sub_rn(1.f, w1y);
    px = min(px, w - 2); py = min(py, h - 2);
    const int t = py * w + px;
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v00 = texel[t * 3 + c], v10 = texel[(t + 1) * 3 + c], v01 = texel[(t + w) * 3 + c], v11 = texel[(t + w + 1) * 3 + c];
        const float a = fma_rn(w0x, v00, mul_rn(w1x, v10)), b = fma_rn(w0x, v01, mul_rn(w1x, v11));
        o[c] = fma_rn(w0y, a, mul_rn(w1y, b));
    }
    pmf[k] = mul_rn(add_rn(add_rn(mul_rn(o[0], .2126f), mul_rn(o[1], .7152f)), mul_rn(o[2], .0722f)), sin_theta[j]);
}
void launch_envmap_pmf(cudaStream_t st, int rx, int ry, int w, int h, const float *texel, const float *sin_theta, float *pmf) {
    k_envmap_pmf<<<nblk((long long)rx * ry, 256), 256, 0, st>>>(rx, ry, w, h, texel, sin_theta, pmf);
}

// ---- scene bounds over the triangle table (min / max are order-independent) -----------------------------------------------------
__global__ void __launch_bounds__(256) k_tri_bounds(int n, const TriRec *__restrict__ tri, float *__restrict__ lohi) {   // lohi: 6 floats, initialised to +-FLT_MAX
    __shared__ float s_lo[3][8], s_hi[3][8];
    float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const float4 *q = reinterpret_cast<const float4 *>(tri + t);
        const float4 p0 = ldg4(q), e1 = ldg4(q + 1), e2 = ldg4(q + 2);
        const float a[3] = {p0.x, p0.y, p0.z}, b[3] = {add_rn(p0.x, e1.x), add_rn(p0.y, e1.y), add_rn(p0.z, e1.z)}, c[3] = {add_rn(p0.x, e2.x), add_rn(p0.y, e2.y), add_rn(p0.z, e2.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], fminf(a[k], fminf(b[k], c[k]))); hi[k] = fmaxf(hi[k], fmaxf(a[k], fmaxf(b[k], c[k]))); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; ++k) { s_lo[k][threadIdx.x >> 5] = lo[k]; s_hi[k][threadIdx.x >> 5] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float l = s_lo[threadIdx.x][0], h = s_hi[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) { l = fminf(l, s_lo[threadIdx.x][w]); h = fmaxf(h, s_hi[threadIdx.x][w]); }
        // float atomics on ordered ints: positive / negative floats order like / opposite to their bit patterns
        auto amin = [](float *addr, float v) { if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v)); else atomicMax(reinterpret_cast<unsigned *>(addr), __float_as_uint(v)); };
        auto amax = [](float *addr, float v) { if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v)); else atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(v)); };
        amin(lohi + threadIdx.x, l); amax(lohi + 3 + threadIdx.x, h);
    }
}
void launch_tri_bounds(cudaStream_t st, int n, const TriRec *tri, float *lohi) {
    const float init[6] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    cudaMemcpyAsync(lohi, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (n > 0) k_tri_bounds<<<std::min(nblk(n, 256), 592u), 256, 0, st>>>(n, tri, lohi);
}

}  // namespace pb
