// fp64 3-vector algebra with a FIXED operation order (compiled with -fmad=false).
// dot(a,b) = a0*b0 + (a1*b1 + a2*b2) — the reduction tree Eigen uses for fixed-size-3 vectors,
// i.e. what Drake's Vector3d::dot evaluates.
#pragma once
#include <cuda_runtime.h>

namespace hcs {

struct D3 {
	double x, y, z;
};
__host__ __device__ __forceinline__ D3 mk(double x, double y, double z) { return D3{ x, y, z }; }
__host__ __device__ __forceinline__ D3 operator+(D3 a, D3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ D3 operator-(D3 a, D3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ D3 operator-(D3 a) { return mk(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ D3 operator*(D3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ D3 operator*(double s, D3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ D3 operator/(D3 a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__host__ __device__ __forceinline__ D3 cross(D3 a, D3 b)
{
	return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ D3 normalized(D3 a)
{
	double z = dot(a, a);
	return z > 0 ? a / sqrt(z) : a;
}
__host__ __device__ __forceinline__ D3 ld3(const double *p) { return mk(p[0], p[1], p[2]); }

// Four doubles fetched with ONE 256-bit load (LDG.E.256 on sm_100a; p must be 32-byte aligned).  The gather kernels
// read per-lane records (every lane another tet / triangle): an 8-byte load touches 32 different lines and costs
// 32 L1 wavefronts per instruction whatever its width, so the records are laid out in 32-byte groups and read whole.
struct __align__(32) D4 {
	double x, y, z, w;
};
__device__ __forceinline__ D4 ld4(const double *p) { return *reinterpret_cast<const D4 *>(p); }
// Records that are read once per step (alive queries: records.cuh ld8f_stream; candidates: __ldcs) are loaded without L1
// allocation, so that the records every lane gathers again and again (tet fields, triangles, tree nodes: a few tens of KB
// per geom) stay in the L1 that the kernels' shared memory leaves (HCS_STREAM_LOADS=0: plain loads).
#ifndef HCS_STREAM_LOADS
#define HCS_STREAM_LOADS 1
#endif
__host__ __device__ __forceinline__ D3 xyz(const D4 &q) { return mk(q.x, q.y, q.z); }

struct Xform { // p_A = R * p_B + p, R row-major
	double R[9];
	D3 p;
};
__host__ __device__ __forceinline__ D3 rot(const double *R, D3 v)
{
	return mk(dot(mk(R[0], R[1], R[2]), v), dot(mk(R[3], R[4], R[5]), v), dot(mk(R[6], R[7], R[8]), v));
}
__host__ __device__ __forceinline__ D3 rotT(const double *R, D3 v)
{
	return mk(dot(mk(R[0], R[3], R[6]), v), dot(mk(R[1], R[4], R[7]), v), dot(mk(R[2], R[5], R[8]), v));
}
__host__ __device__ __forceinline__ D3 apply(const Xform &X, D3 v) { return rot(X.R, v) + X.p; }
// RigidTransform::InvertAndCompose: X_AC = X_BA^-1 * X_BC
__host__ __device__ __forceinline__ Xform invert_and_compose(const Xform &BA, const Xform &BC)
{
	Xform X;
#pragma unroll
	for (int i = 0; i < 3; ++i)
#pragma unroll
		for (int j = 0; j < 3; ++j)
			X.R[3 * i + j] = dot(mk(BA.R[i], BA.R[3 + i], BA.R[6 + i]), mk(BC.R[j], BC.R[3 + j], BC.R[6 + j]));
	X.p = rotT(BA.R, BC.p - BA.p);
	return X;
}
__device__ __forceinline__ Xform load_pose(const double *xpos, const double *xmat, int n_geoms, int env, int g)
{
	Xform X;
	const double *m = xmat + ((size_t)env * n_geoms + g) * 9;
#pragma unroll
	for (int i = 0; i < 9; ++i)
		X.R[i] = m[i];
	X.p = ld3(xpos + ((size_t)env * n_geoms + g) * 3);
	return X;
}

// cos(5*pi/8), Drake's kAlpha cull threshold (glibc value of std::cos(5.*M_PI/8.))
#define HCS_COS_ALPHA (-0x1.87de2a6aea962p-2)

} // namespace hcs
