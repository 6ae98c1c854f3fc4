// K1: se(3) knots -> P camera-to-world poses on a cubic B-spline (or a line) in one launch.
//
// Replaces spline.py:247-331 + model/optimize.py:58-111, which issue O(300) tiny ATen
// launches per call.  One thread per timestamp; the four knot conversions are recomputed by
// every thread (about 400 flops) instead of being staged, which keeps the kernel free of
// synchronisation.  The arithmetic is written on a scalar template so that the same code
// runs on forward-mode dual numbers for the backward pass (d poses / d knots).
//
// Built with -fmad=false: the reference evaluates every product and sum as a separately
// rounded fp32 op, and pose errors are amplified 512x by the positional encoding.
#include "common.cuh"

namespace bnrf {

template <class T> struct Quat { T x, y, z, w; };
template <class T> struct Vec3 { T x, y, z; };

__device__ inline float s_sqrt(float a) { return sqrtf(a); }
__device__ inline float s_sin(float a) { return sinf(a); }
__device__ inline float s_cos(float a) { return cosf(a); }
__device__ inline float s_atan(float a) { return atanf(a); }
__device__ inline float s_val(float a) { return a; }
__device__ inline float s_powi(float a, int n) {       // torch.pow(x, n): n=0 -> 1, n=2 -> x*x, else powf
    if (n == 0) return 1.0f;
    if (n == 2) return a * a;
    return powf(a, (float)n);
}

// Forward-mode dual number: value and one directional derivative.  interpolate_pose<Dual> seeded on knot element j
// yields d pose / d knot_j, which spline_backward_kernel contracts with d L / d pose.
struct Dual {
    float v, d;
    __device__ Dual() : v(0.f), d(0.f) {}
    __device__ Dual(float a) : v(a), d(0.f) {}
    __device__ Dual(float a, float b) : v(a), d(b) {}
};
__device__ inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ inline Dual operator/(Dual a, Dual b) { const float q = a.v / b.v; return {q, (a.d - q * b.d) / b.v}; }
__device__ inline Dual s_sqrt(Dual a) { const float r = sqrtf(a.v); return {r, a.d * 0.5f / r}; }
__device__ inline Dual s_sin(Dual a) { return {sinf(a.v), cosf(a.v) * a.d}; }
__device__ inline Dual s_cos(Dual a) { return {cosf(a.v), -sinf(a.v) * a.d}; }
__device__ inline Dual s_atan(Dual a) { return {atanf(a.v), a.d / (1.0f + a.v * a.v)}; }
__device__ inline float s_val(Dual a) { return a.v; }
__device__ inline Dual s_powi(Dual a, int n) {
    if (n == 0) return Dual(1.0f);
    if (n == 2) return a * a;
    return {powf(a.v, (float)n), (float)n * powf(a.v, (float)(n - 1)) * a.d};
}

// sum_i (-1)^i x^(2i) / d_i,  d_i = prod_{j<=i} (2j+first)(2j+first+1)   (spline.py:46-62)
template <class T>
__device__ T taylor_series(T x, int first) {
    T total = T(0.0f);
    double den = 1.0;
#pragma unroll 1
    for (int i = 0; i <= 10; ++i) {
        den *= (double)((2 * i + first) * (2 * i + first + 1));
        T term = s_powi(x, 2 * i) / T((float)den);
        total = (i & 1) ? total - term : total + term;
    }
    return total;
}

// The backward kernel's instance: the same series with x^(2i) built by repeated multiplication instead of powf (two powf per term,
// 176 per pose and knot element, were most of the kernel's 78 us of dependent latency).  The forward kernel keeps powf: it is
// compared with the reference's poses at fp32 rounding, the derivatives only at the gradient tolerance.
template <>
__device__ Dual taylor_series<Dual>(Dual x, int first) {
    Dual total(0.0f);
    double den = 1.0;
    const float x2 = x.v * x.v;
    float p = 1.0f, pm1 = 0.0f;            // x^(2i), x^(2i-1)
#pragma unroll 1
    for (int i = 0; i <= 10; ++i) {
        den *= (double)((2 * i + first) * (2 * i + first + 1));
        const float inv = 1.0f / (float)den;
        const Dual term(p * inv, (float)(2 * i) * pm1 * x.d * inv);
        total = (i & 1) ? total - term : total + term;
        pm1 = p * x.v;
        p *= x2;
    }
    return total;
}

// exp map, both regimes (spline.py:79-100); the series branch is selected for half-angle < 1e-9
template <class T>
__device__ Quat<T> rotvec_to_quat(Vec3<T> r) {
    T half = T(0.5f) * s_sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
    Quat<T> q;
    if (s_val(half) < 1e-9f) {
        T h2 = half * half, h4 = h2 * h2;
        T k = T(0.5f) - T(1.0f / 12.0f) * h2 - T(1.0f / 240.0f) * h4;
        q.x = k * r.x; q.y = k * r.y; q.z = k * r.z;
        q.w = T(1.0f) - T(0.5f) * h2 + T(1.0f / 24.0f) * h4;
    } else {
        T lam = s_sin(half) / (T(2.0f) * half);
        q.x = lam * r.x; q.y = lam * r.y; q.z = lam * r.z; q.w = s_cos(half);
    }
    return q;
}

// log map with atan (not atan2) and its |w| ~ 0 / theta ~ 0 regimes (spline.py:167-192)
template <class T>
__device__ Vec3<T> quat_to_rotvec(Quat<T> q) {
    T th = s_sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    T lam;
    const float w = s_val(q.w);
    if (fabsf(w) < 1e-10f) {
        lam = (w < 0.0f) ? T(-3.14159265358979323846f) / th : T(3.14159265358979323846f) / th;
    } else if (s_val(th) < 1e-20f) {
        lam = T(2.0f) / q.w - T(2.0f / 3.0f) * (th * th) / (q.w * q.w * q.w);
    } else {
        lam = T(2.0f) * s_atan(th / q.w) / th;
    }
    return {lam * q.x, lam * q.y, lam * q.z};
}

// a (x) b through the left-multiplication matrix of spline.py:130-138
template <class T>
__device__ Quat<T> quat_mul(Quat<T> a, Quat<T> b) {
    Quat<T> r;
    r.x = a.w * b.x + (T(0.0f) - a.z) * b.y + a.y * b.z + a.x * b.w;
    r.y = a.z * b.x + a.w * b.y + (T(0.0f) - a.x) * b.z + a.y * b.w;
    r.z = (T(0.0f) - a.y) * b.x + a.x * b.y + a.w * b.z + a.z * b.w;                        // row [-y, x, w, z]
    r.w = (T(0.0f) - a.x) * b.x + (T(0.0f) - a.y) * b.y + (T(0.0f) - a.z) * b.z + a.w * b.w;  // row [-x,-y,-z, w]
    return r;
}
template <class T>
__device__ Quat<T> quat_conj(Quat<T> q) { return {T(0.0f) - q.x, T(0.0f) - q.y, T(0.0f) - q.z, q.w}; }

// knot (w, u) -> quaternion and translation V(w) u  (spline.py:16-26)
template <class T>
__device__ void knot_to_qt(const T* k, Quat<T>& q, Vec3<T>& t) {
    Vec3<T> w{k[0], k[1], k[2]}, u{k[3], k[4], k[5]};
    T th = s_sqrt(w.x * w.x + w.y * w.y + w.z * w.z);
    T B = taylor_series(th, 1), C = taylor_series(th, 2);
    // [w]x u and [w]x^2 u, with [w]x^2 = w w^T - |w|^2 I expanded entry by entry as a 3x3 product
    T m[3][3] = {{T(0.0f), T(0.0f) - w.z, w.y}, {w.z, T(0.0f), T(0.0f) - w.x}, {T(0.0f) - w.y, w.x, T(0.0f)}};
    T uu[3] = {u.x, u.y, u.z}, out[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T acc = T(0.0f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // python precedence: C * wx @ wx == (C * wx) @ wx
            T m2 = (C * m[r][0]) * m[0][c] + (C * m[r][1]) * m[1][c] + (C * m[r][2]) * m[2][c];
            T v = ((r == c) ? T(1.0f) : T(0.0f)) + B * m[r][c] + m2;
            acc = acc + v * uu[c];
        }
        out[r] = acc;
    }
    t = {out[0], out[1], out[2]};
    q = rotvec_to_quat(w);
}

template <class T>
__device__ void quat_to_pose(Quat<T> q, Vec3<T> t, T* P) {   // spline.py:111-118, row-major [3,4]
    T b = q.x, c = q.y, d = q.z, a = q.w;
    P[0] = T(1.0f) - T(2.0f) * (c * c + d * d); P[1] = T(2.0f) * (b * c - a * d); P[2] = T(2.0f) * (a * c + b * d); P[3] = t.x;
    P[4] = T(2.0f) * (b * c + a * d); P[5] = T(1.0f) - T(2.0f) * (b * b + d * d); P[6] = T(2.0f) * (c * d - a * b); P[7] = t.y;
    P[8] = T(2.0f) * (b * d - a * c); P[9] = T(2.0f) * (a * b + c * d); P[10] = T(1.0f) - T(2.0f) * (b * b + c * c); P[11] = t.z;
}

// One interpolated pose.  knots: 4 x 6 values (transform already added); traj 0 cubic, 1 linear.
template <class T>
__device__ void interpolate_pose(const T* knots, float time, int traj, T* P) {
    // u == 0 -> 1e-6, u == 1 -> 1 - 1e-6 (spline.py:249-252)
    if (time == 0.0f) time = time + 0.000001f;
    if (time == 1.0f) time = time - 0.000001f;
    const float u = time;
    Quat<T> q0, q3; Vec3<T> t0, t3;
    knot_to_qt(knots + 0, q0, t0);
    knot_to_qt(knots + 18, q3, t3);
    if (traj == 1) {   // spline.py:305-331
        Vec3<T> r = quat_to_rotvec(quat_mul(quat_conj(q0), q3));
        Quat<T> q = quat_mul(q0, rotvec_to_quat<T>({T(u) * r.x, T(u) * r.y, T(u) * r.z}));
        const float a = 1.0f - u;
        quat_to_pose<T>(q, {T(a) * t0.x + T(u) * t3.x, T(a) * t0.y + T(u) * t3.y, T(a) * t0.z + T(u) * t3.z}, P);
        return;
    }
    Quat<T> q1, q2; Vec3<T> t1, t2;
    knot_to_qt(knots + 6, q1, t1);
    knot_to_qt(knots + 12, q2, t2);
    const float uu = u * u, uuu = u * u * u, s6 = (float)(1.0 / 6.0), h = 0.5f;
    // translation: uniform cubic B-spline basis (spline.py:267-273)
    const float c0 = s6 - h * u + h * uu - s6 * uuu;
    const float c1 = 4.0f * s6 - uu + h * uuu;
    const float c2 = s6 + h * u + h * uu - h * uuu;
    const float c3 = s6 * uuu;
    Vec3<T> t{T(c0) * t0.x + T(c1) * t1.x + T(c2) * t2.x + T(c3) * t3.x,
              T(c0) * t0.y + T(c1) * t1.y + T(c2) * t2.y + T(c3) * t3.y,
              T(c0) * t0.z + T(c1) * t1.z + T(c2) * t2.z + T(c3) * t3.z};
    // rotation: cumulative basis on the relative-rotation logs (spline.py:276-297)
    const float b1 = (float)(5.0 / 6.0) + h * u - h * uu + s6 * uuu;
    const float b2 = s6 + h * u + h * uu - 2.0f * s6 * uuu;
    const float b3 = s6 * uuu;
    Vec3<T> r01 = quat_to_rotvec(quat_mul(quat_conj(q0), q1));
    Vec3<T> r12 = quat_to_rotvec(quat_mul(quat_conj(q1), q2));
    Vec3<T> r23 = quat_to_rotvec(quat_mul(quat_conj(q2), q3));
    Quat<T> e0 = rotvec_to_quat<T>({r01.x * T(b1), r01.y * T(b1), r01.z * T(b1)});
    Quat<T> e1 = rotvec_to_quat<T>({r12.x * T(b2), r12.y * T(b2), r12.z * T(b2)});
    Quat<T> e2 = rotvec_to_quat<T>({r23.x * T(b3), r23.y * T(b3), r23.z * T(b3)});
    Quat<T> q = quat_mul(q0, quat_mul(e0, quat_mul(e1, e2)));
    quat_to_pose(q, t, P);
}

// poses p < n_plain interpolate the knots as they are (event camera), the others knots + transform (RGB camera)
__global__ void spline_kernel(const float* __restrict__ knots, const float* __restrict__ transform,
                              const float* __restrict__ ts, int P, int n_plain, int traj, float* __restrict__ poses) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    if (p < n_plain) transform = nullptr;
    float k[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) k[i] = knots[i] + (transform ? transform[i % 6] : 0.0f);   // optimize.py:86-89
    if (!transform) {
#pragma unroll
        for (int i = 0; i < 24; ++i) k[i] = knots[i];
    }
    float out[12];
    interpolate_pose<float>(k, ts[p], traj, out);
#pragma unroll
    for (int i = 0; i < 12; ++i) poses[p * 12 + i] = out[i];
}

// d L / d knots [4,6] (+= ) and d L / d transform [6] (+=) from d L / d poses [P,3,4]: thread (p, j) differentiates
// pose p along knot element j.  The RGB knots are knots + transform (optimize.py:86-89), so the transform's
// gradient is the sum of the four knots' gradients element-wise.
__global__ void spline_backward_kernel(const float* __restrict__ knots, const float* __restrict__ transform,
                                       const float* __restrict__ ts, int P, int n_plain, int traj, const float* __restrict__ d_poses,
                                       float* __restrict__ d_knots, float* __restrict__ d_transform) {
    const int p = blockIdx.x, j = threadIdx.x;
    if (p >= P || j >= 24) return;
    if (p < n_plain) transform = nullptr;
    Dual k[24];
    for (int i = 0; i < 24; ++i) k[i] = Dual(knots[i] + (transform ? transform[i % 6] : 0.0f), i == j ? 1.0f : 0.0f);
    Dual out[12];
    interpolate_pose<Dual>(k, ts[p], traj, out);
    float g = 0.f;
    for (int i = 0; i < 12; ++i) g += out[i].d * d_poses[p * 12 + i];
    if (g != g) g = 0.f;                       // 0/0 of a derivative at an exactly-zero rotation (the reference yields NaN there)
    atomicAdd(d_knots + j, g);
    if (transform && d_transform) atomicAdd(d_transform + (j % 6), g);
}

int launch_spline_backward(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int n_plain, int traj,
                           const float* d_poses, float* d_knots, float* d_transform, cudaStream_t st) {
    if (!knots || !ts || !d_poses || !d_knots || P <= 0 || n_plain < 0 || n_plain > P || (traj != 0 && traj != 1))
        return fail(ctx, BNRF_ERR_ARG, "spline_backward: bad argument");
    spline_backward_kernel<<<P, 32, 0, st>>>(knots, transform, ts, P, n_plain, traj, d_poses, d_knots, d_transform);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

int launch_spline(bnrf_ctx* ctx, const float* knots, const float* transform, const float* ts, int P, int n_plain, int traj,
                  float* poses, cudaStream_t st) {
    if (!knots || !ts || !poses || P <= 0 || n_plain < 0 || n_plain > P || (traj != 0 && traj != 1)) return fail(ctx, BNRF_ERR_ARG, "spline_poses: bad argument");
    spline_kernel<<<(P + 63) / 64, 64, 0, st>>>(knots, transform, ts, P, n_plain, traj, poses);
    BNRF_LAUNCH_CHECK(ctx);
    return BNRF_OK;
}

}  // namespace bnrf
