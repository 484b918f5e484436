/*
 * pn2_oracle.c -- CPU restatement of the reference lib/pointnet2 CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under situation3d_b200/ may link, import
 * or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline -- never as the product path.
 *
 * Each function follows one reference kernel (paths relative to
 * /root/reference/lib/pointnet2/_ext_src/) and reproduces its floating-point
 * recipe exactly: nvcc contracts every a*a + b*b + c*c expression into
 * FMUL, FFMA, FFMA (x term first), so the restatement uses fmaf() in that
 * order and is compiled with -ffp-contract=off so that gcc adds no
 * contraction of its own.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 4).  The oracle is pinned against the reference's own CUDA
 * extension, built unmodified from /root/reference into oracle/_ref/ by
 * oracle/build_ref.py and executed on the GPU box (tests/test_ops_gpu.py::test_against_reference_cuda_kernels_full_size),
 * and against the fixtures those runs produced (tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PN2O_API __attribute__((visibility("default")))

/* cuda_utils.h:13-19  opt_n_threads(): 2^floor(log2 w) clamped to [1, 512].
 * Same double expression as the reference so that any libm rounding quirk is
 * shared. */
PN2O_API int pn2o_opt_n_threads(int work_size)
{
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 512) v = 512;
    if (v < 1) v = 1;
    return v;
}

PN2O_API int pn2o_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    /* (a-b)*(a-b) + ... contracted by nvcc as FMUL, FFMA, FFMA */
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* Same expression as contracted in three_nn_kernel: there nvcc (12.9, -O3, sm_100) multiplies the
 * y term first and fuses the x term into it (SASS of oracle/_ref/interpolate_gpu.cuda.o:
 * FMUL dy*dy; FFMA dx*dx + .; FFMA dz*dz + .).  Both are legal contractions of the source
 * expression; the golden vectors from the reference build (tests/golden/ref_cuda_ops.npz) pin it. */
static inline float sqdist3_yxz(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------
 * Furthest point sampling.
 * sampling_gpu.cu:69-173 (kernel), :175-229 (block-size dispatch),
 * sampling.cpp:70-76 (temp initialised to 1e10, output zero-filled).
 * The block structure is emulated literally -- block_size virtual threads,
 * strided scan with strict '>' inside a thread, then the shared-memory tree
 * whose strict '>' makes the lower slot win ties -- because the tie order is
 * an artefact of that structure.
 * ------------------------------------------------------------------------ */
static void fps_one(int n, int m, const float *xyz, float *temp, int *idxs)
{
    if (m <= 0) return;
    const int bs = pn2o_opt_n_threads(n);
    float *dists = (float *)malloc(sizeof(float) * (size_t)bs);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)bs);
    int old = 0;
    idxs[0] = old;
    for (int j = 1; j < m; ++j) {
        const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
        for (int tid = 0; tid < bs; ++tid) {
            int besti = 0;
            float best = -1.0f;
            for (int k = tid; k < n; k += bs) {
                const float x2 = xyz[k * 3 + 0], y2 = xyz[k * 3 + 1], z2 = xyz[k * 3 + 2];
                const float mag = fmaf(z2, z2, fmaf(y2, y2, x2 * x2));
                if ((double)mag <= 1e-3) continue;   /* float promoted against a double literal */
                const float d = sqdist3(x2, y2, z2, x1, y1, z1);
                const float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                if (d2 > best) { besti = k; best = d2; }
            }
            dists[tid] = best;
            dists_i[tid] = besti;
        }
        for (int s = bs / 2; s >= 1; s >>= 1) {
            for (int tid = 0; tid < s; ++tid) {
                const float v1 = dists[tid], v2 = dists[tid + s];
                const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                dists[tid] = fmaxf(v1, v2);
                dists_i[tid] = v2 > v1 ? i2 : i1;
            }
        }
        old = dists_i[0];
        idxs[j] = old;
    }
    free(dists);
    free(dists_i);
}

PN2O_API void pn2o_furthest_point_sampling(int b, int n, int m, const float *xyz, int *idxs)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) {
        float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        for (int k = 0; k < n; ++k) temp[k] = 1e10f;
        for (int j = 0; j < m; ++j) idxs[(size_t)i * m + j] = 0;
        if (n > 0) fps_one(n, m, xyz + (size_t)i * n * 3, temp, idxs + (size_t)i * m);
        free(temp);
    }
}

/* ------------------------------------------------------------------------
 * Ball query.  ball_query_gpu.cu:9-44; zero-filled output ball_query.cpp:19-21.
 * First nsample indices in ascending order with d2 < r*r (strict), padded
 * with the first hit.
 * ------------------------------------------------------------------------ */
PN2O_API void pn2o_ball_query(int b, int n, int m, float radius, int nsample,
                              const float *new_xyz, const float *xyz, int *idx)
{
    const float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < m; ++j) {
            const float *q = new_xyz + ((size_t)i * m + j) * 3;
            const float *p = xyz + (size_t)i * n * 3;
            int *out = idx + ((size_t)i * m + j) * nsample;
            for (int l = 0; l < nsample; ++l) out[l] = 0;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sqdist3(q[0], q[1], q[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) out[l] = k;
                    out[cnt] = k;
                    ++cnt;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * three_nn.  interpolate_gpu.cu:9-59.  Running bests are doubles initialised
 * to 1e40, compared with strict '<' against the float distance; output is
 * the squared distance (the sqrt lives in pointnet2_utils.py:142).
 * ------------------------------------------------------------------------ */
PN2O_API void pn2o_three_nn(int b, int n, int m, const float *unknown, const float *known,
                            float *dist2, int *idx)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < n; ++j) {
            const float *u = unknown + ((size_t)i * n + j) * 3;
            const float *kn = known + (size_t)i * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int b1 = 0, b2 = 0, b3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = sqdist3_yxz(u[0], u[1], u[2], kn[k * 3], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; b3 = b2;
                    best2 = best1; b2 = b1;
                    best1 = d; b1 = k;
                } else if (d < best2) {
                    best3 = best2; b3 = b2;
                    best2 = d; b2 = k;
                } else if (d < best3) {
                    best3 = d; b3 = k;
                }
            }
            float *od = dist2 + ((size_t)i * n + j) * 3;
            int *oi = idx + ((size_t)i * n + j) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = b1; oi[1] = b2; oi[2] = b3;
        }
    }
}

/* three_interpolate.  interpolate_gpu.cu:72-101:  p1*w1 + p2*w2 + p3*w3
 * contracted as FMUL (second term), FFMA (first term), FFMA (third term) in the reference
 * build -- pinned by tests/golden/ref_cuda_ops.npz. */
PN2O_API void pn2o_three_interpolate(int b, int c, int m, int n, const float *points,
                                     const int *idx, const float *weight, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i) {
        for (int l = 0; l < c; ++l) {
            const float *pl = points + ((size_t)i * c + l) * m;
            float *ol = out + ((size_t)i * c + l) * n;
            for (int j = 0; j < n; ++j) {
                const int *id = idx + ((size_t)i * n + j) * 3;
                const float *w = weight + ((size_t)i * n + j) * 3;
                ol[j] = fmaf(pl[id[2]], w[2], fmaf(pl[id[0]], w[0], pl[id[1]] * w[1]));
            }
        }
    }
}

/* three_interpolate_grad.  interpolate_gpu.cu:116-143.  The reference
 * accumulates with float atomicAdd in a nondeterministic order; the oracle
 * adds in ascending (j, t) order into a zero-filled output. */
PN2O_API void pn2o_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                          const int *idx, const float *weight,
                                          float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i) {
        for (int l = 0; l < c; ++l) {
            const float *gl = grad_out + ((size_t)i * c + l) * n;
            float *ol = grad_points + ((size_t)i * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const int *id = idx + ((size_t)i * n + j) * 3;
                const float *w = weight + ((size_t)i * n + j) * 3;
                ol[id[0]] += gl[j] * w[0];
                ol[id[1]] += gl[j] * w[1];
                ol[id[2]] += gl[j] * w[2];
            }
        }
    }
}

/* gather_points.  sampling_gpu.cu:8-20. */
PN2O_API void pn2o_gather_points(int b, int c, int n, int m, const float *points,
                                 const int *idx, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                out[((size_t)i * c + l) * m + j] =
                    points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}

/* gather_points_grad.  sampling_gpu.cu:34-47 (atomicAdd scatter). */
PN2O_API void pn2o_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                      const int *idx, float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] +=
                    grad_out[((size_t)i * c + l) * m + j];
}

/* group_points.  group_points_gpu.cu:8-28. */
PN2O_API void pn2o_group_points(int b, int c, int n, int npoints, int nsample,
                                const float *points, const int *idx, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i) {
        for (int l = 0; l < c; ++l) {
            const float *pl = points + ((size_t)i * c + l) * n;
            const int *id = idx + (size_t)i * npoints * nsample;
            float *ol = out + ((size_t)i * c + l) * npoints * nsample;
            for (size_t t = 0; t < (size_t)npoints * nsample; ++t) ol[t] = pl[id[t]];
        }
    }
}

/* group_points_grad.  group_points_gpu.cu:43-64 (atomicAdd scatter). */
PN2O_API void pn2o_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_out, const int *idx,
                                     float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < b; ++i) {
        for (int l = 0; l < c; ++l) {
            const float *gl = grad_out + ((size_t)i * c + l) * npoints * nsample;
            const int *id = idx + (size_t)i * npoints * nsample;
            float *ol = grad_points + ((size_t)i * c + l) * n;
            for (size_t t = 0; t < (size_t)npoints * nsample; ++t) ol[id[t]] += gl[t];
        }
    }
}

/* ---- point <-> pixel correspondences -----------------------------------------------
 * lib/projection.py, ProjectionHelper.  The reference evaluates its 3- and 4-term dot products with torch.mm /
 * torch.bmm (summation order left to the BLAS); this restatement fixes them as fmaf chains in index order and
 * keeps every element-wise step separately rounded, which is also what csrc/projection.cu computes.  Pinned
 * against the reference's own class run on the CPU (tests/golden/make_ref_projection_goldens.py): index lists
 * identical on the golden scenes, corner / normal values to fp32 rounding. */

/* compute_frustum_corners (:48-70) and compute_frustum_normals (:72-119).
 * corner_points (8,3): _compute_corner_points (:29-46), w = 1.  corners (8,4), normals (6,3). */
PN2O_API void pn2o_frustum_planes(const float *c2w, const float *corner_points, float *corners, float *normals)
{
    for (int k = 0; k < 8; ++k)
        for (int r = 0; r < 4; ++r) {
            float acc = c2w[r * 4 + 0] * corner_points[k * 3 + 0];
            acc = fmaf(c2w[r * 4 + 1], corner_points[k * 3 + 1], acc);
            acc = fmaf(c2w[r * 4 + 2], corner_points[k * 3 + 2], acc);
            acc = fmaf(c2w[r * 4 + 3], 1.0f, acc);
            corners[k * 4 + r] = acc;
        }
    /* plane k: origin corner, end of plane_vec1, end of plane_vec2 (front, right, roof, left, bottom, back) */
    static const int tri[6][3] = {{0, 3, 1}, {1, 2, 5}, {2, 3, 6}, {3, 0, 7}, {0, 1, 4}, {5, 6, 4}};
    for (int k = 0; k < 6; ++k) {
        float a[3], b[3];
        for (int j = 0; j < 3; ++j) {
            a[j] = corners[tri[k][1] * 4 + j] - corners[tri[k][0] * 4 + j];
            b[j] = corners[tri[k][2] * 4 + j] - corners[tri[k][0] * 4 + j];
        }
        const float p0 = a[1] * b[2], q0 = a[2] * b[1], p1 = a[2] * b[0], q1 = a[0] * b[2], p2 = a[0] * b[1], q2 = a[1] * b[0];
        normals[k * 3 + 0] = p0 - q0;
        normals[k * 3 + 1] = p1 - q1;
        normals[k * 3 + 2] = p2 - q2;
    }
}

/* compute_projection (:191-254) for one view.  intr4 = {fx, fy, cx, cy}, range3 = {depth_min, depth_max, accuracy}.
 * idx3d / idx2d: (n+1) int64, element 0 = count; returns the count (0 where the reference returns None). */
PN2O_API long long pn2o_compute_projection(int n, const float *points, const float *depth, const float *c2w,
                                           const float *w2c, const float *intr4, const float *range3, int width,
                                           int height, const float *corner_points, long long *idx3d, long long *idx2d)
{
    float corners[32], normals[18];
    pn2o_frustum_planes(c2w, corner_points, corners, normals);
    memset(idx3d, 0, sizeof(long long) * ((size_t)n + 1));
    memset(idx2d, 0, sizeof(long long) * ((size_t)n + 1));
    long long cnt = 0;
    for (int i = 0; i < n; ++i) {
        const float x = points[(size_t)i * 3], y = points[(size_t)i * 3 + 1], z = points[(size_t)i * 3 + 2];
        /* points_in_frustum (:139-150): planes 0-2 measured from corner 2, planes 3-5 from corner 4 */
        int in = 1;
        for (int k = 0; k < 6; ++k) {
            const float *o = corners + (k < 3 ? 2 : 4) * 4;
            const float px = x - o[0], py = y - o[1], pz = z - o[2];
            const float d = fmaf(pz, normals[k * 3 + 2], fmaf(py, normals[k * 3 + 1], px * normals[k * 3 + 0]));
            const float r = rintf(d * 100.0f);                 /* torch.round: half to even */
            if (!(r < 0.0f)) in = 0;                            /* round(..) / 100 < 0 */
        }
        if (!in) continue;
        /* world -> camera (:223) */
        float cam[3];
        for (int r = 0; r < 3; ++r) {
            float acc = w2c[r * 4 + 0] * x;
            acc = fmaf(w2c[r * 4 + 1], y, acc);
            acc = fmaf(w2c[r * 4 + 2], z, acc);
            cam[r] = fmaf(w2c[r * 4 + 3], 1.0f, acc);
        }
        /* camera -> image, rounded to a pixel (:226-228) */
        const float mu = cam[0] * intr4[0], mv = cam[1] * intr4[1];
        const float du = mu / cam[2], dv = mv / cam[2];
        const float u = rintf(du + intr4[2]), v = rintf(dv + intr4[3]);
        if (!(u >= 0.0f && v >= 0.0f && u < (float)width && v < (float)height)) continue;      /* :231 */
        const long long pix = (long long)v * width + (long long)u;                               /* :236 */
        const float dval = depth[pix];
        const float diff = dval - cam[2];
        if (!(dval >= range3[0] && dval <= range3[1] && fabsf(diff) <= range3[2])) continue;     /* :239-240 */
        idx3d[1 + cnt] = i;
        idx2d[1 + cnt] = pix;
        ++cnt;
    }
    idx3d[0] = cnt;
    idx2d[0] = cnt;
    return cnt;
}

/* points_in_frustum (lib/projection.py:121-155) for given corners (8,4) and normals (6,3): mask (n) bytes, returns
 * the number of points inside. */
PN2O_API long long pn2o_points_in_frustum(int n, const float *points, const float *corners, const float *normals,
                                          unsigned char *mask)
{
    long long cnt = 0;
    for (int i = 0; i < n; ++i) {
        const float x = points[(size_t)i * 3], y = points[(size_t)i * 3 + 1], z = points[(size_t)i * 3 + 2];
        int in = 1;
        for (int k = 0; k < 6; ++k) {
            const float *o = corners + (k < 3 ? 2 : 4) * 4;
            const float px = x - o[0], py = y - o[1], pz = z - o[2];
            const float d = fmaf(pz, normals[k * 3 + 2], fmaf(py, normals[k * 3 + 1], px * normals[k * 3 + 0]));
            if (!(rintf(d * 100.0f) < 0.0f)) in = 0;
        }
        if (mask) mask[i] = (unsigned char)in;
        cnt += in;
    }
    return cnt;
}
