// 256-bit gathers of the resident geometry records (hcs_internal.h): one LDG.E.256 per 32-byte group.
// Every lane of a gather kernel reads another record, so each load instruction costs 32 L1 wavefronts whatever
// its width; reading a 192-byte TetField with 6 loads instead of 24 is what matters (profiles/r01_notes.md).
#pragma once
#include "dmath.cuh"
#include "hcs_internal.h"

namespace hcs {

struct TetVerts {
	D3 v0, v1, v2, v3;
	__device__ __forceinline__ D3 at(int i) const { return i == 0 ? v0 : (i == 1 ? v1 : (i == 2 ? v2 : v3)); }
};
__device__ __forceinline__ TetVerts load_tet_verts(const TetGeom *g) // v[4][3] = three groups
{
	const double *p = reinterpret_cast<const double *>(g);
	D4 a = ld4(p), b = ld4(p + 4), c = ld4(p + 8);
	TetVerts t;
	t.v0 = mk(a.x, a.y, a.z), t.v1 = mk(a.w, b.x, b.y), t.v2 = mk(b.z, b.w, c.x), t.v3 = mk(c.y, c.z, c.w);
	return t;
}
__device__ __forceinline__ D4 load_tet_pressures(const TetGeom *g) // e[4] = the fourth group
{
	return ld4(reinterpret_cast<const double *>(g) + 12);
}
__device__ __forceinline__ double pick(const D4 &q, int i) { return i == 0 ? q.x : (i == 1 ? q.y : (i == 2 ? q.z : q.w)); }

struct TriVerts {
	D3 v0, v1, v2, n;
};
__device__ __forceinline__ TriVerts load_tri(const TriRec *r) // v[3][3] + n[3] = three groups
{
	const double *p = reinterpret_cast<const double *>(r);
	D4 a = ld4(p), b = ld4(p + 4), c = ld4(p + 8);
	TriVerts t;
	t.v0 = mk(a.x, a.y, a.z), t.v1 = mk(a.w, b.x, b.y), t.v2 = mk(b.z, b.w, c.x), t.n = mk(c.y, c.z, c.w);
	return t;
}

// TetLeaf32: four 32-byte groups of eight floats: planes 0-1 | planes 2-3 | ghat, -, v0, v1.x | v1.yz, v2, v3
struct __align__(32) F8 {
	float a[8];
};
__device__ __forceinline__ F8 ld8f(const TetLeaf32 *t, int group) { return reinterpret_cast<const F8 *>(t)[group]; }
// eight floats that are read once per step (alive records): no L1 allocation (dmath.cuh HCS_STREAM_LOADS)
__device__ __forceinline__ F8 ld8f_stream(const F8 *p)
{
#if HCS_STREAM_LOADS
	F8 q;
	asm("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	    : "=f"(q.a[0]), "=f"(q.a[1]), "=f"(q.a[2]), "=f"(q.a[3]), "=f"(q.a[4]), "=f"(q.a[5]), "=f"(q.a[6]), "=f"(q.a[7])
	    : "l"(p));
	return q;
#else
	return *p;
#endif
}
// TetLeafSS32: grad, e0, ghat, - | v0, v1, v2.xy | v2.z, v3, -
__device__ __forceinline__ F8 ld8f(const TetLeafSS32 *t, int group) { return reinterpret_cast<const F8 *>(t)[group]; }
__device__ __forceinline__ float fdot3(float ax, float ay, float az, float bx, float by, float bz)
{
	return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, az * bz));
}

// TetField groups: plane[0..3] (unit normal + offset), {grad, e0}, {ghat, -}
__device__ __forceinline__ D4 load_plane(const TetField *f, int k) { return ld4(reinterpret_cast<const double *>(f) + 4 * k); }
__device__ __forceinline__ D4 load_grad_e0(const TetField *f) { return ld4(reinterpret_cast<const double *>(f) + 16); }
__device__ __forceinline__ D3 load_ghat(const TetField *f) { return xyz(ld4(reinterpret_cast<const double *>(f) + 20)); }

} // namespace hcs
