// voxel_pe.cu -- per-axis positional embedding looked up by integer voxel coordinate.
//
// Reference: 3DLLM_BLIP2-base/lavis/models/blip2_models/blip2_t5.py:104-118 (and :279-293, the same loop in
// predict_answers) and blip2_opt.py:92-104.  Per sample j the reference indexes a (256, 469) sinusoid table with
// the x, y and z voxel coordinate of every point, concatenates the three 469-wide rows into channels 0..1406 of
// a zero (P, 1408) CPU tensor, moves it to the GPU and then
//   T5 site :  pc_embeds = pc_embeds + 0.01 * all_pcs                      blip2_t5.py:118
//   OPT site:  pc_embeds = torch.cat([pc_embeds, all_pcs], 1)              blip2_opt.py:104
// Here that is one HBM-bound pass: a warp owns a point, loads its three coordinates once, and streams the
// point's feature row through registers as float4 while the table rows come from L1/L2 (the table is 480 KB and
// is the only thing read more than once).  Algorithmic bytes per point: c*4 read + c*4 written (add mode).
//
// Arithmetic is the reference's: the product 0.01f * pe is rounded to fp32 before the add (two roundings, no FMA),
// so the result is bit-identical to torch's.  Index semantics are torch's: `.long()` truncates a float coordinate
// toward zero, a negative index counts from the end of the table, anything else is an IndexError in the reference
// -- here *status is set to 1 (the row is clamped so that nothing is read out of bounds) and the caller raises.
#include "common.cuh"

namespace pn2 {

constexpr int kPeThreads = 256;
constexpr int kPeWarps = kPeThreads / kWarp;

template <typename CT> __device__ __forceinline__ long long pe_coord(CT v);
template <> __device__ __forceinline__ long long pe_coord<long long>(long long v) { return v; }
template <> __device__ __forceinline__ long long pe_coord<int>(int v) { return v; }
template <> __device__ __forceinline__ long long pe_coord<float>(float v)      // cvt.rzi, as .long(); NaN / inf -> out of range
{
    return fabsf(v) < 1e18f ? (long long)v : (long long)1 << 62;
}

template <int VEC> struct PeVec;
template <> struct PeVec<4> {
    float x[4];
    __device__ __forceinline__ void load(const float *p, int v) { const float4 t = __ldcs(reinterpret_cast<const float4 *>(p) + v); x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w; }
    __device__ __forceinline__ void store(float *p, int v) const { __stcs(reinterpret_cast<float4 *>(p) + v, make_float4(x[0], x[1], x[2], x[3])); }
};
template <> struct PeVec<1> {
    float x[1];
    __device__ __forceinline__ void load(const float *p, int v) { x[0] = __ldcs(p + v); }
    __device__ __forceinline__ void store(float *p, int v) const { __stcs(p + v, x[0]); }
};

constexpr int kPeBatch = 4;      // vectors a lane keeps in flight: 4 x (16 B of features + 4 table words)

// MODE 0: out = feat + scale * pe.  MODE 1: out rows [p, 2p) of each sample = pe, rows [0, p) = feat when COPY.
template <typename CT, int VEC, int MODE, bool COPY>
__global__ void __launch_bounds__(kPeThreads)
voxel_pe_kernel(long long points, int p, int c, int seg, int table_rows, int coord_stride, const CT *__restrict__ coords,
                const float *__restrict__ table, const float *__restrict__ feat, float *__restrict__ out, float scale,
                int *__restrict__ status)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kPeWarps + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kPeWarps;
    const int nvec = c / VEC;
    for (long long pt = warp0; pt < points; pt += nwarps) {
        // the point's three table rows
        int row[3];
        bool bad = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            long long i = pe_coord<CT>(coords[pt * coord_stride + a]);
            if (i < 0) i += table_rows;
            if (i < 0 || i >= table_rows) { bad = true; i = 0; }
            row[a] = (int)i * seg;
        }
        if (bad && lane == 0) atomicExch(status, 1);
        // channel ch of axis a reads table[row[a] + ch - a * seg]: fold the segment offset into the row offset
        const int o0 = row[0], o1 = row[1] - seg, o2 = row[2] - 2 * seg;
        const float *src = feat + pt * c;
        float *dst_pe, *dst_copy = nullptr;
        if (MODE == 0) {
            dst_pe = out + pt * c;
        } else {                                  // cat along the point axis: sample s owns rows [2ps, 2ps + 2p)
            const long long s = pt / p, j = pt - s * p;
            dst_copy = out + (2 * s * p + j) * c;
            dst_pe = out + ((2 * s + 1) * p + j) * c;
        }
        for (int v0 = lane; v0 < nvec; v0 += 32 * kPeBatch) {
            PeVec<VEC> pe[kPeBatch], f[kPeBatch];
            // every load of the batch first (selects and predicated loads, no branches in between) ...
#pragma unroll
            for (int u = 0; u < kPeBatch; ++u) {
                const int v = v0 + 32 * u;
                const bool live = v < nvec;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int ch = v * VEC + e;
                    const int off = ch < seg ? o0 : ch < 2 * seg ? o1 : o2;
                    pe[u].x[e] = (live && ch < 3 * seg) ? __ldg(table + off + ch) : 0.f;
                }
                if ((MODE == 0 || COPY) && live) f[u].load(src, v);
            }
            // ... then the arithmetic and the stores
#pragma unroll
            for (int u = 0; u < kPeBatch; ++u) {
                const int v = v0 + 32 * u;
                if (v >= nvec) break;
                if (MODE == 0) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pe[u].x[e] = __fadd_rn(f[u].x[e], __fmul_rn(scale, pe[u].x[e]));
                } else if (COPY) {
                    f[u].store(dst_copy, v);
                }
                pe[u].store(dst_pe, v);
            }
        }
    }
}

template <typename CT, int VEC>
static void launch_voxel_pe_mode(int grid, cudaStream_t stream, long long points, int p, int c, int seg, int table_rows,
                                 int coord_stride, const CT *coords, const float *table, const float *feat, float *out,
                                 float scale, int mode, int *status)
{
    if (mode == PN2_VOXEL_PE_ADD)
        voxel_pe_kernel<CT, VEC, 0, false><<<grid, kPeThreads, 0, stream>>>(points, p, c, seg, table_rows, coord_stride, coords, table, feat, out, scale, status);
    else if (feat)
        voxel_pe_kernel<CT, VEC, 1, true><<<grid, kPeThreads, 0, stream>>>(points, p, c, seg, table_rows, coord_stride, coords, table, feat, out, scale, status);
    else
        voxel_pe_kernel<CT, VEC, 1, false><<<grid, kPeThreads, 0, stream>>>(points, p, c, seg, table_rows, coord_stride, coords, table, feat, out, scale, status);
}

template <typename CT>
static int launch_voxel_pe(long long points, int p, int c, int seg, int table_rows, int coord_stride, const void *coords,
                           const float *table, const float *feat, float *out, float scale, int mode, int *status,
                           cudaStream_t stream)
{
    const bool vec4 = c % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const long long want = (points + kPeWarps - 1) / kPeWarps;
    const long long cap = (long long)stream_sm_count(stream) * 8;          // grid-stride over points beyond that
    const int grid = (int)(want < cap ? want : cap);
    if (vec4)
        launch_voxel_pe_mode<CT, 4>(grid, stream, points, p, c, seg, table_rows, coord_stride, static_cast<const CT *>(coords), table, feat, out, scale, mode, status);
    else
        launch_voxel_pe_mode<CT, 1>(grid, stream, points, p, c, seg, table_rows, coord_stride, static_cast<const CT *>(coords), table, feat, out, scale, mode, status);
    PN2_LAUNCH_CHECK("voxel_pe");
    return PN2_OK;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_voxel_pe(int b, int p, int c, int seg, int table_rows, int coord_kind, int coord_stride, const void *coords,
                            const float *table, const float *feat, float *out, float scale, int mode, int *status,
                            pn2_stream_t stream)
{
    if (b < 0 || p < 0 || c < 1 || seg < 1 || table_rows < 1 || coord_stride < 3 || (long long)3 * seg > c) return PN2_ERR_INVALID_ARGUMENT;
    if (mode != PN2_VOXEL_PE_ADD && mode != PN2_VOXEL_PE_CAT) return PN2_ERR_INVALID_ARGUMENT;
    if ((long long)table_rows * seg > 0x7fffffffLL) return PN2_ERR_INVALID_ARGUMENT;
    if (!status) return PN2_ERR_INVALID_ARGUMENT;
    PN2_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    if (b == 0 || p == 0) return PN2_OK;
    if (!coords || !table || !out || (mode == PN2_VOXEL_PE_ADD && !feat)) return PN2_ERR_INVALID_ARGUMENT;
    const long long points = (long long)b * p;
    switch (coord_kind) {
    case PN2_COORD_I32:
        return launch_voxel_pe<int>(points, p, c, seg, table_rows, coord_stride, coords, table, feat, out, scale, mode, status, as_stream(stream));
    case PN2_COORD_I64:
        return launch_voxel_pe<long long>(points, p, c, seg, table_rows, coord_stride, coords, table, feat, out, scale, mode, status, as_stream(stream));
    case PN2_COORD_F32:
        return launch_voxel_pe<float>(points, p, c, seg, table_rows, coord_stride, coords, table, feat, out, scale, mode, status, as_stream(stream));
    }
    return PN2_ERR_INVALID_ARGUMENT;
}
