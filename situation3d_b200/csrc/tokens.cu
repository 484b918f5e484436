// tokens.cu -- visual-token construction in front of the re-encoding.
//
// Reference: situation3d/models/sqa_module.py:297-315, a per-scene Python loop over the sparse encoder's
// bottleneck voxels (coords (m,3) int, feats (m,c) f32):
//   reduced = coords[:, [0, 1]];  unique_coords, indices = reduced.unique(dim=0, return_inverse=True)     :298-299
//   reduced_feats = zeros(U, c).scatter_reduce_(0, indices..., feats, reduce='mean')                        :300-301
//   sampled = randperm(U)[:T]   (or randperm(U) ++ randint(0, U, T - U) when U < T)                         :303-308
//   tokens = reduced_feats[sampled];  positions = (unique_coords[sampled] + stride[0:2] / 2) * voxel_size   :309-311
// Two details of that code decide the numbers and are reproduced exactly: torch.unique(dim=0) returns the
// columns in ascending (x, y) order, and scatter_reduce_('mean') onto a zero tensor counts the zero
// (include_self): a column of n voxels gets  sum / (n + 1), not the mean.
//
// pn2_column_pool: one CTA per scene.  A 64-bit key (x, y, voxel index) per voxel is sorted in shared memory
// (bitonic), column heads are flagged and scanned, and one warp per column sums its voxels' features in
// ascending voxel order (the order a sequential scatter visits them) -- coalesced over channels.
// pn2_token_gather: the sampled columns (indices drawn by the caller with the reference's RNG calls) ->
// (b, t, c) tokens and (b, t, 2) positions in metres.
#include "common.cuh"

namespace pn2 {

constexpr int kColThreads = 512;
constexpr int kColMax = 8192;            // voxels per scene (13-bit voxel index in the sort key)
constexpr int kCoordBias = 1 << 19;      // |x|, |y| < 2^19

__global__ void __launch_bounds__(kColThreads)
column_pool_kernel(const int *__restrict__ offsets, const int *__restrict__ coords, const float *__restrict__ feats, int c,
                   int *__restrict__ ncols, int *__restrict__ out_coords, float *__restrict__ out_feats,
                   int *__restrict__ inverse, int *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char col_smem[];
    unsigned long long *key = reinterpret_cast<unsigned long long *>(col_smem);      // [n2]
    int *col = reinterpret_cast<int *>(key + kColMax);                                // [n2] column of sorted position i
    int *start = col + kColMax;                                                       // [ncols + 1] first sorted position
    __shared__ int part[kColThreads];
    __shared__ int bad;

    const int s = blockIdx.x, tid = threadIdx.x;
    const int m0 = offsets[s], m = offsets[s + 1] - m0;
    if (tid == 0) bad = 0;
    __syncthreads();
    if (m <= 0) { if (tid == 0) ncols[s] = 0; return; }
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = tid; i < n2; i += kColThreads) {
        unsigned long long k = ~0ull;
        if (i < m) {
            const int x = coords[(size_t)(m0 + i) * 3], y = coords[(size_t)(m0 + i) * 3 + 1];
            if (x < -kCoordBias || x >= kCoordBias || y < -kCoordBias || y >= kCoordBias) bad = 1;
            k = ((unsigned long long)(unsigned)(x + kCoordBias) << 33) | ((unsigned long long)(unsigned)(y + kCoordBias) << 13) |
                (unsigned long long)i;
        }
        key[i] = k;
    }
    __syncthreads();
    if (bad) { if (tid == 0) { ncols[s] = 0; atomicExch(status, 1); } return; }
    // bitonic sort, ascending: (x, y) lexicographic as signed integers, then voxel index
    for (int k2 = 2; k2 <= n2; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += kColThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = key[i], b = key[p];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { key[i] = b; key[p] = a; }
                }
            }
            __syncthreads();
        }
    // column heads -> column numbers (block-wide scan: 16 consecutive positions per thread)
    const int per = n2 / kColThreads > 0 ? n2 / kColThreads : 1;
    const int lo = tid * per, hi = min(lo + per, m);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += (i == 0 || (key[i] >> 13) != (key[i - 1] >> 13)) ? 1 : 0;
    part[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < kColThreads; ++t) { const int v = part[t]; part[t] = run; run += v; }
        ncols[s] = run;
        start[run] = m;
    }
    __syncthreads();
    int run = part[tid];
    for (int i = lo; i < hi; ++i) {
        const bool head = i == 0 || (key[i] >> 13) != (key[i - 1] >> 13);
        if (head) {
            start[run] = i;
            out_coords[(size_t)(m0 + run) * 2] = (int)(unsigned)(key[i] >> 33) - kCoordBias;
            out_coords[(size_t)(m0 + run) * 2 + 1] = (int)(unsigned)((key[i] >> 13) & 0xFFFFFu) - kCoordBias;
            ++run;
        }
        col[i] = run - 1;
        if (inverse) inverse[m0 + (int)(key[i] & (kColMax - 1))] = run - 1;
    }
    __syncthreads();
    // one warp per column: sum of its voxels' features in ascending voxel order, divided by (count + 1)
    const int u = ncols[s], warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < u; q += kColThreads / 32) {
        const int a = start[q], b = start[q + 1];
        const float denom = (float)(b - a + 1);
        for (int ch = lane; ch < c; ch += 32) {
            float sum = 0.f;
            for (int i = a; i < b; ++i) sum = __fadd_rn(sum, __ldg(feats + (size_t)(m0 + (int)(key[i] & (kColMax - 1))) * c + ch));
            out_feats[(size_t)(m0 + q) * c + ch] = __fdiv_rn(sum, denom);
        }
    }
}

__global__ void token_gather_kernel(int t, int c, const int *__restrict__ offsets, const int *__restrict__ sampled,
                                    const float *__restrict__ pooled, const int *__restrict__ pooled_coords, float half_x,
                                    float half_y, float voxel_size, float *__restrict__ tokens, float *__restrict__ positions)
{
    const int s = blockIdx.y, j = blockIdx.x;
    const int src = offsets[s] + sampled[(size_t)s * t + j];
    const float *row = pooled + (size_t)src * c;
    float *dst = tokens + ((size_t)s * t + j) * c;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) dst[ch] = __ldg(row + ch);
    if (threadIdx.x == 0) {
        // (coords + stride / 2) * voxel_size, fp32 as torch evaluates it (sqa_module.py:311)
        positions[((size_t)s * t + j) * 2] = __fmul_rn(__fadd_rn((float)pooled_coords[(size_t)src * 2], half_x), voxel_size);
        positions[((size_t)s * t + j) * 2 + 1] = __fmul_rn(__fadd_rn((float)pooled_coords[(size_t)src * 2 + 1], half_y), voxel_size);
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_column_pool_max_voxels(void) { return kColMax; }

extern "C" int pn2_column_pool(int b, const int *offsets, const int *coords, const float *feats, int c, int *ncols,
                               int *out_coords, float *out_feats, int *inverse, int *status, pn2_stream_t stream)
{
    if (b < 0 || c < 1) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0) return PN2_OK;
    if (!offsets || !coords || !feats || !ncols || !out_coords || !out_feats || !status) return PN2_ERR_INVALID_ARGUMENT;
    const size_t smem = (size_t)kColMax * 8 + (size_t)kColMax * 4 + (size_t)(kColMax + 1) * 4;
    PN2_CUDA_TRY(cudaFuncSetAttribute(column_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PN2_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    column_pool_kernel<<<b, kColThreads, smem, as_stream(stream)>>>(offsets, coords, feats, c, ncols, out_coords, out_feats,
                                                                     inverse, status);
    PN2_LAUNCH_CHECK("column_pool");
    return PN2_OK;
}

extern "C" int pn2_token_gather(int b, int t, int c, const int *offsets, const int *sampled, const float *pooled,
                                const int *pooled_coords, float half_x, float half_y, float voxel_size, float *tokens,
                                float *positions, pn2_stream_t stream)
{
    if (b < 0 || t < 0 || c < 1) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || t == 0) return PN2_OK;
    if (!offsets || !sampled || !pooled || !pooled_coords || !tokens || !positions || b > 65535) return PN2_ERR_INVALID_ARGUMENT;
    token_gather_kernel<<<dim3(t, b), 128, 0, as_stream(stream)>>>(t, c, offsets, sampled, pooled, pooled_coords, half_x, half_y,
                                                                   voxel_size, tokens, positions);
    PN2_LAUNCH_CHECK("token_gather");
    return PN2_OK;
}
