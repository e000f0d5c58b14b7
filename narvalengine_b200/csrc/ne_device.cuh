// ne_device.cuh — device functions of the hot path. Each function names the reference code whose behaviour it
// reproduces (paths relative to /root/reference/src). The estimator is reproduced WITH its quirks (SURVEY.md
// Appendix A, Q1..Q30) because image parity is judged against the reference's own CPU render.
//
// Everything is templated on the uniform source R (TapeRng: reference draw order for tape-exact tests;
// PhiloxRng: production) and on two compile-time switches:
//   FAITHFUL  consume uniforms / do work whose result the reference discards (visibilityTr's ratio tracking
//             behind `return 0`, intersectTr in medium-free scenes) so a tape stays aligned draw for draw.
//   BRICKMAJ  track with per-brick majorants over a brick DDA instead of the reference's single global majorant
//             (same expectation, far fewer null collisions; SURVEY.md A.5 last bullet).
#pragma once
#include <climits>

#include "ne_math.cuh"
#include "ne_rng.cuh"
#include "ne_scene.cuh"

namespace ne {

struct Stats {
	uint32_t extend_rays, shadow_rays, delta_steps, ratio_steps, brick_visits, bvh_nodes, tri_tests, prim_tests, scatter_events,
		surface_events;
	NE_D void clear() {
		extend_rays = shadow_rays = delta_steps = ratio_steps = brick_visits = bvh_nodes = tri_tests = prim_tests = scatter_events =
			surface_events = 0;
	}
};

// ---------------------------------------------------------------------------------------------------------------
// Textures: Texture::sample / sampleAtIndex / wrapTextureCoordinates, materials/Texture.cpp:37-129;
// Material::sampleMaterial, materials/Material.h:62-69
// ---------------------------------------------------------------------------------------------------------------
NE_D V4 tex_at(const DTexture& t, uint32_t index) {
	V4 r = {0, 0, 0, 0};
	if (t.format == TEX_RGBA8) {
		uchar4 c = reinterpret_cast<const uchar4*>(t.texels)[index];
		r.x = c.x / 255.0f; r.y = c.y / 255.0f; r.z = c.z / 255.0f; r.w = c.w / 255.0f;
		return r;
	}
	const float* p = reinterpret_cast<const float*>(t.texels);
	switch (t.format) {
	case TEX_R32F: r.x = __ldg(p + index); break;
	case TEX_RG32F: r.x = __ldg(p + 2 * index); r.y = __ldg(p + 2 * index + 1); break;
	case TEX_RGB32F: r.x = __ldg(p + 3 * index); r.y = __ldg(p + 3 * index + 1); r.z = __ldg(p + 3 * index + 2); break;
	default: r.x = __ldg(p + 4 * index); r.y = __ldg(p + 4 * index + 1); r.z = __ldg(p + 4 * index + 2); r.w = __ldg(p + 4 * index + 3); break;
	}
	return r;
}
NE_D float tex_wrap(float u, int mode) {
	if (mode == 1) return gclamp(fabsf(u - float(int(u))), 0.0f, 1.0f);  // "mirror"
	return gclamp(u, 0.0f, 1.0f);
}
NE_D V4 tex_sample(const DTexture& t, float u, float v) {
	u = tex_wrap(u, t.wrap_u);
	v = tex_wrap(v, t.wrap_v);
	int x = int(u * t.w), y = int(v * t.h);
	if (x > 0 && x == t.w) x--;
	if (y > 0 && y == t.h) y--;
	return tex_at(t, uint32_t(t.w * y + x));
}
NE_D V4 material_sample(const DScene& s, int tex, float u, float v) {
	if (tex < 0) { V4 r = {0, 0, 0, 1}; return r; }
	return tex_sample(s.tex[tex], u, v);
}

// ---------------------------------------------------------------------------------------------------------------
// InfiniteAreaLight (lights/InfiniteAreaLight.h:20-131) + Distribution1D/2D (utils/Sampling.h:5-112, Q27)
// ---------------------------------------------------------------------------------------------------------------
// binarySearch, utils/Math.h:1128-1146 (pbrt FindInterval): last index with cdf[index] <= u, clamped to [0, size-2]
NE_D int dist_find(const float* cdf, int size, float u) {
	int first = 0, len = size;
	while (len > 0) {
		int half = len >> 1, middle = first + half;
		if (__ldg(cdf + middle) <= u) { first = middle + 1; len -= half + 1; }
		else len = half;
	}
	return min(max(first - 1, 0), size - 2);
}
// Distribution1D::sampleContinuous, utils/Sampling.h:35-52
NE_D float dist1d_sample(const float* func, const float* cdf, float funcInt, int n, float u, float& pdf, int& off) {
	off = dist_find(cdf, n + 1, u);
	float c0 = __ldg(cdf + off), c1 = __ldg(cdf + off + 1);
	float du = u - c0;
	if ((c1 - c0) > 0) du /= (c1 - c0);
	pdf = (funcInt > 0) ? __ldg(func + off) / funcInt : 0.0f;
	return (float(off) + du) / float(n);
}
// Distribution2D::sampleContinuous :96-104: u[1] picks the row, u[0] the column within that row's conditional
NE_D void env_sample_continuous(const DEnvDist& e, float u0, float u1, float& d0, float& d1, float& pdf) {
	float p0, p1;
	int v, dummy;
	d1 = dist1d_sample(e.mFunc, e.mCdf, e.mInt, e.h, u1, p1, v);
	d0 = dist1d_sample(e.cFunc + size_t(v) * e.nc, e.cCdf + size_t(v) * (e.nc + 1), __ldg(e.cInt + v), e.nc, u0, p0, dummy);
	pdf = p0 * p1;
}
// InfiniteAreaLight::Le :47-56 for a WCS direction: w = normalize(invM * d), local frame of normal (0,0,1)
NE_D V3 env_le(const DScene& s, const DEnvDist& e, const float* invM, V3 dW) {
	V3 w = normalize(xform_dir(invM, dW));
	V3 normal(0.0f, 0.0f, 1.0f), ss, ts;
	onb(normal, ss, ts);
	V3 in = to_lcs(w, normal, ss, ts);
	float p = atan2f(in.y, in.x);
	float phi = (p < 0) ? float(double(p) + NE_TWO_PI) : p;
	float theta = acosf(gclamp(in.z, -1.0f, 1.0f));
	float su = float(double(phi) * 0.15915494309189533577), sv = float(double(theta) * 0.31830988618379067154);
	V4 t = tex_sample(s.tex[e.tex], su, sv);
	return V3(t.x, t.y, t.z);
}
NE_D V3 env_le_identity(const DScene& s, const DEnvDist& e, V3 dW) {
	const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	return env_le(s, e, I, dW);
}

// ---------------------------------------------------------------------------------------------------------------
// Primitives
// ---------------------------------------------------------------------------------------------------------------
// AABB::intersect, primitives/AABB.cpp:48-79 (centre/half-size slab; true even when the box is behind the ray;
// miss -> tNear=+inf, tFar=-inf).
NE_D bool aabb_intersect(V3 bmin, V3 bmax, Ray r, float& tNear, float& tFar, V3& normal, V3& hitPoint) {
	V3 s((r.d.x < 0.0f) ? 1.0f : -1.0f, (r.d.y < 0.0f) ? 1.0f : -1.0f, (r.d.z < 0.0f) ? 1.0f : -1.0f);
	V3 invD = 1.0f / r.d;
	V3 size = bmax - bmin;
	V3 half = size / 2.0f;
	V3 center = bmax - size / 2.0f;
	V3 t1 = (center - r.o + s * half) * invD;
	V3 t2 = (center - r.o - s * half) * invD;
	V3 tmn = gmin(t1, t2), tmx = gmax(t1, t2);
	float tn = gmax(gmax(tmn.x, tmn.y), tmn.z);
	float tf = gmin(gmin(tmx.x, tmx.y), tmx.z);
	V3 sg(gsign(r.d.x), gsign(r.d.y), gsign(r.d.z));
	normal = -sg * V3(gstep(t1.y, t1.x), gstep(t1.z, t1.y), gstep(t1.x, t1.z)) * V3(gstep(t1.z, t1.x), gstep(t1.x, t1.y), gstep(t1.y, t1.z));
	tNear = tn;
	tFar = tf;
	hitPoint = r.at(tn);
	if (tn > tf) {
		tNear = INFINITY;
		tFar = -INFINITY;
		return false;
	}
	return true;
}

// Rectangle::barycentricCoordinates + samplePointOnTexture, primitives/Rectangle.cpp:90-157, for the unit square
// SceneReader builds (SceneReader.cpp:407-468): vertices a,b,c,d = (-.5,-.5,0),(.5,-.5,0),(.5,.5,0),(-.5,.5,0),
// uv (0,0),(1,0),(1,1),(0,1).
NE_D void bary3(V3 p, V3 a, V3 b, V3 c, float& l0, float& l1, float& l2) {
	V3 v0 = b - a, v1 = c - a, v2 = p - a;
	float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
	float denom = d00 * d11 - d01 * d01;
	l1 = (d11 * d20 - d01 * d21) / denom;
	l2 = (d00 * d21 - d01 * d20) / denom;
	l0 = 1.0f - l1 - l2;
}
NE_D void rect_uv(V3 p, float& u, float& v) {
	const V3 a(-0.5f, -0.5f, 0), b(0.5f, -0.5f, 0), c(0.5f, 0.5f, 0), d(-0.5f, 0.5f, 0);
	float x0, x1, x2, y0, y1, y2;
	bary3(p, a, b, c, x0, x1, x2);
	bary3(p, d, b, c, y0, y1, y2);
	if (x0 < 0 || x1 < 0 || x2 < 0 || x0 > 1 || x1 > 1 || x2 > 1) {
		// w*D + u*B + v*C with uv D=(0,1), B=(1,0), C=(1,1)
		u = y0 * 0.0f + y1 * 1.0f + y2 * 1.0f;
		v = y0 * 1.0f + y1 * 0.0f + y2 * 1.0f;
	} else {
		u = x0 * 0.0f + x1 * 1.0f + x2 * 1.0f;
		v = x0 * 0.0f + x1 * 0.0f + x2 * 1.0f;
	}
}
// Rectangle::intersect, primitives/Rectangle.cpp:50-76 (plane z=0, normal (0,0,-1), corners (-.5,-.5,0),(.5,.5,0)).
NE_D bool rect_intersect(Ray r, Hit& hit) {
	const V3 normal(0.0f, 0.0f, -1.0f);
	const V3 planeVertex(0.5f, 0.5f, 0.0f);
	float denom = dot(normal, r.d);
	if (double(fabsf(denom)) < NE_EPSILON) return false;
	float num = dot(normal, planeVertex - r.o);
	float t = num / denom;
	if (t < 0) return false;
	V3 p = r.at(t);
	// containsPoint with size (1,1,0): only x and y are tested
	if (p.x < -0.5f || p.x > 0.5f) return false;
	if (p.y < -0.5f || p.y > 0.5f) return false;
	hit.tNear = t;
	hit.tFar = t;
	hit.n = normal;
	hit.p = p;
	hit.prim = 0;
	rect_uv(p, hit.u, hit.v);
	return true;
}
// Sphere::intersect, primitives/Sphere.cpp:21-44 (centre = origin of the OCS; true for ANY real roots).
NE_D bool sphere_intersect(Ray r, float radius, Hit& hit) {
	V3 center(0.0f, 0.0f, 0.0f);
	V3 oc = r.o - center;
	float a = dot(r.d, r.d);
	float b = dot(oc, r.d);
	float c = dot(oc, oc) - radius * radius;
	float disc = b * b - a * c;
	if (disc >= 0) {
		// `sqrt(discriminant)` resolves to ::sqrt(double) in the reference build: roots are formed in double, then narrowed
		double sq = sqrt(double(disc)), nb = double(-b), ad = double(a);
		float t1 = float((nb - sq) / ad), t2 = float((nb + sq) / ad);
		hit.tNear = gmin(t1, t2);
		hit.tFar = gmax(t1, t2);
		hit.p = r.at(hit.tNear);
		hit.n = normalize((hit.p - center) / radius);
		hit.u = hit.v = 0;
		hit.prim = 0;
		return true;
	}
	return false;
}

// Triangle::intersect, primitives/Triangle.cpp:49-81 — the t/u/v part. Accepts t >= 0, u,v in [0,1], u+v <= 1.
NE_D bool tri_intersect(V3 v0, V3 v1, V3 v2, Ray ray, float& tOut) {
	V3 v1v0 = v1 - v0, v2v0 = v2 - v0, rov0 = ray.o - v0;
	V3 n = cross(v1v0, v2v0);
	V3 q = cross(rov0, ray.d);
	float denom = 1.0f / dot(ray.d, n);
	float u = dot(-q, v2v0) * denom;
	float v = dot(q, v1v0) * denom;
	float t = dot(-n, rov0) * denom;
	if (isnan(t) || t < 0) return false;
	if (u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f) return false;
	tOut = t;
	return true;
}

// convertNormalFromTextureMap, utils/Math.h:1209-1215
NE_D V3 normal_from_map(V3 texNormal, V3 worldNormal) {
	V3 nt = texNormal * 2.0f - 1.0f;
	V3 t = cross(worldNormal, V3(0.0f, 1.0f, 0.0f));
	V3 b = normalize(cross(worldNormal, t));
	// mat3(t,b,n) * nt = t*nt.x + b*nt.y + n*nt.z
	V3 r = t * nt.x + b * nt.y + worldNormal * nt.z;
	return normalize(r);
}

// Closest triangle with t >= 0 (BVH::intersect semantics, primitives/BVH.cpp:108-194: no tMin/tMax inside, strict
// `<` keeps the first of equal hits) over OUR binned-SAH 2-wide BVH. Node boxes are tested with a conservative
// slab test; every triangle test uses the reference arithmetic above.
NE_D V3 bvh_inv_dir(const Ray& r) {
	// A zero direction component would give 0*inf = NaN for rays lying exactly in a box face; a huge finite reciprocal
	// keeps the slab test inclusive there (the reference's one-triangle unit test hits at a vertex, tests.cpp:300-347).
	return V3(1.0f / (fabsf(r.d.x) > 1e-20f ? r.d.x : copysignf(1e-20f, r.d.x)), 1.0f / (fabsf(r.d.y) > 1e-20f ? r.d.y : copysignf(1e-20f, r.d.y)),
	          1.0f / (fabsf(r.d.z) > 1e-20f ? r.d.z : copysignf(1e-20f, r.d.z)));
}
#define NE_BVH_DONE INT_MIN  // bottom of the traversal stack
#define NE_BVH_STACK 48
// One inner-node visit: both child boxes against the ray, near child first, far child pushed.
NE_D int bvh_node_step(const DMesh& m, const Ray& r, const V3& inv, int node, int* stack, int& sp, float tBest) {
	const float4* np = reinterpret_cast<const float4*>(m.nodes + node);
	float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
	// n0 = lo0.xyz hi0.x ; n1 = hi0.yz lo1.xy ; n2 = lo1.z hi1.xyz ; n3 = child0 child1 - -
	float ax = (n0.x - r.o.x) * inv.x, bx = (n0.w - r.o.x) * inv.x;
	float ay = (n0.y - r.o.y) * inv.y, by = (n1.x - r.o.y) * inv.y;
	float az = (n0.z - r.o.z) * inv.z, bz = (n1.y - r.o.z) * inv.z;
	float tn0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
	float tf0 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
	ax = (n1.z - r.o.x) * inv.x; bx = (n2.y - r.o.x) * inv.x;
	ay = (n1.w - r.o.y) * inv.y; by = (n2.z - r.o.y) * inv.y;
	az = (n2.x - r.o.z) * inv.z; bz = (n2.w - r.o.z) * inv.z;
	float tn1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
	float tf1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.0000004f;
	bool h0 = tn0 <= tf0 && tn0 <= tBest;
	bool h1 = tn1 <= tf1 && tn1 <= tBest;
	int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
	if (h0 && h1) {
		if (tn1 < tn0) { int t = c0; c0 = c1; c1 = t; }
		if (sp < NE_BVH_STACK) stack[sp++] = c1;
		return c0;
	}
	if (h0) return c0;
	if (h1) return c1;
	return stack[--sp];
}
// One leaf (node < 0, not NE_BVH_DONE): its triangles with the reference arithmetic, strict `<` keeps the first of equals.
NE_D void bvh_leaf(const DMesh& m, const Ray& r, int node, float& tBest, int& slotBest, bool& any, Stats& st) {
	int enc = ~node;
	int first = enc >> 3, cnt = enc & 7;
	for (int i = 0; i < cnt; i++) {
		const float4* tp = m.tri + 3 * size_t(first + i);
		float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
		float t;
		st.tri_tests++;
		if (tri_intersect(V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), r, t)) {
			any = true;
			if (t < tBest) { tBest = t; slotBest = first + i; }
		}
	}
}
NE_D bool bvh_closest(const DMesh& m, Ray r, float& tBest, int& slotBest, Stats& st) {
	const V3 inv = bvh_inv_dir(r);
	// "while-while" traversal (Aila & Laine 2009): every lane first descends inner nodes until it stands on a leaf (or
	// has nothing left), and only then do the lanes of the warp test their leaves' triangles together. With the two kinds
	// of work interleaved per lane ("if-if"), ncu showed 3 of 32 lanes active in the triangle code on the 2 M-triangle
	// scene. The order in which a ray visits nodes and triangles - hence the result - is unchanged.
	int stack[NE_BVH_STACK];
	int sp = 0;
	stack[sp++] = NE_BVH_DONE;
	int node = 0;
	tBest = INFINITY;
	slotBest = -1;
	bool any = false;
	while (true) {
		while (node >= 0) {
			st.bvh_nodes++;
			node = bvh_node_step(m, r, inv, node, stack, sp, tBest);
		}
		if (node == NE_BVH_DONE) break;
		bvh_leaf(m, r, node, tBest, slotBest, any, st);
		node = stack[--sp];
	}
	return any;
}

// Triangle::samplePointOnTexture, primitives/Triangle.cpp:121-147 with Q30: the three vertices/uvs are read
// CONSECUTIVELY from the vertex buffer starting at the triangle's first vertex (indices i0, i0+1, i0+2).
NE_D void tri_uv(const DMesh& m, int tri, V3 p, float& u, float& v) {
	u = v = 0;
	if (!m.uv) return;
	int i0 = int(m.idx[3 * size_t(tri)]);
	int i1 = min(i0 + 1, m.n_verts - 1), i2 = min(i0 + 2, m.n_verts - 1);
	V3 a(m.pos[3 * i0], m.pos[3 * i0 + 1], m.pos[3 * i0 + 2]);
	V3 b(m.pos[3 * i1], m.pos[3 * i1 + 1], m.pos[3 * i1 + 2]);
	V3 c(m.pos[3 * i2], m.pos[3 * i2 + 1], m.pos[3 * i2 + 2]);
	float l0, l1, l2;
	bary3(p, a, b, c, l0, l1, l2);
	u = l0 * m.uv[2 * i0] + l1 * m.uv[2 * i1] + l2 * m.uv[2 * i2];
	v = l0 * m.uv[2 * i0 + 1] + l1 * m.uv[2 * i1 + 1] + l2 * m.uv[2 * i2 + 1];
}

// ---------------------------------------------------------------------------------------------------------------
// InstancedModel::intersect (primitives/InstancedModel.cpp:24-38) over Model::intersect (primitives/Model.cpp:372-447)
// for the one-primitive models SceneReader builds. `in_lights`: the instance sits in Scene::lights, i.e. its
// primitive is in Model::lights (tested by the lights loop :432-444: no "inside" case).
// ---------------------------------------------------------------------------------------------------------------
// The mesh branch after the BVH has answered (Model.cpp:381-392 + InstancedModel.cpp:30-36): accept the closest triangle
// if it lies in (tMin, tMax), fill the hit record in OCS. `ray` is the OCS ray.
NE_D bool mesh_accept(const DScene& s, const DInstance& in, const DMesh& m, const Ray& ray, bool local, float t, int slot, Hit& hit, float tMin, float& tMax) {
	if (!(local && t > tMin && t < tMax)) return false;
	tMax = t;
	const float4* tp = m.tri + 3 * size_t(slot);
	float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
	V3 v0(a.x, a.y, a.z), v1(b.x, b.y, b.z), v2(c.x, c.y, c.z);
	int tri = __float_as_int(a.w);
	hit.tNear = hit.tFar = t;
	hit.p = ray.at(t);
	tri_uv(m, tri, hit.p, hit.u, hit.v);
	V3 n = normalize(cross(v1 - v0, v2 - v0));
	if (in.material >= 0 && s.mat[in.material].has_normal_flag) {
		V4 tn4 = material_sample(s, s.mat[in.material].normal_tex, hit.u, hit.v);
		n = normal_from_map(V3(tn4.x, tn4.y, tn4.z), n);
	}
	hit.n = n;
	hit.prim = tri;
	return true;
}
// InstancedModel.cpp:33-36: the accepted hit goes back to WCS.
NE_D void hit_to_wcs(const DInstance& in, int i, Hit& hit) {
	hit.p = xform_point(in.M, hit.p);
	hit.n = normalize(xform_dir(in.M, hit.n));
	hit.inst = i;
}

// DEFER = true (the persistent trace kernels, SceneTrace below): a mesh whose bounding box the ray enters is NOT
// traversed here; the call returns false with `deferred` set and the OCS ray in `rayO`, the caller walks the BVH at its
// own pace and completes the instance with mesh_accept + hit_to_wcs.
enum { ISECT_FULL = 0, ISECT_DEFER = 1, ISECT_NOMESH = 2 };  // ISECT_NOMESH: the scene has no mesh (the host checked): no BVH code is compiled in
template <int DEFER>
NE_D bool instance_intersect_t(const DScene& s, int i, Ray rayW, Hit& hit, float tMin, float& tMax, bool in_lights, Stats& st, bool& deferred, Ray& rayO) {
	const DInstance& in = s.inst[i];
	if (!in.collision) return false;
	if (in_lights) {  // Model.cpp:429-432: a light whose Le(ray, identity) is not black (directional, environment) is never intersected
		const DMaterial& lm = s.mat[in.material];
		if (lm.directional && !is_black(V3(lm.li[0], lm.li[1], lm.li[2]))) return false;
		if (lm.infinite && !is_black(env_le_identity(s, s.env[lm.env], rayW.d))) return false;
	}
	Ray ray = transform_ray(rayW, in.Mi);
	bool did = false;
	st.prim_tests++;
	if (in.type == PRIM_MESH) {
		const DMesh& m = s.mesh[in.mesh];
		float tn, tf; V3 nn, hp;
		if (!aabb_intersect(V3(m.bbmin[0], m.bbmin[1], m.bbmin[2]), V3(m.bbmax[0], m.bbmax[1], m.bbmax[2]), ray, tn, tf, nn, hp)) return false;
		if (DEFER == ISECT_NOMESH) return false;
		if (DEFER == ISECT_DEFER) {
			deferred = true;
			rayO = ray;
			return false;
		}
		float t; int slot;
		bool local = bvh_closest(m, ray, t, slot, st);
		did = mesh_accept(s, in, m, ray, local, t, slot, hit, tMin, tMax);
	} else if (in.type == PRIM_VOLUME && !in_lights && s.mat[in.material].volume >= 0) {
		// GridMedia special case, Model.cpp:394-413: ignores tMin/tMax (Q4)
		float tn, tf; V3 nn, hp;
		aabb_intersect(V3(-0.5f, -0.5f, -0.5f), V3(0.5f, 0.5f, 0.5f), ray, tn, tf, nn, hp);
		float a = gmax(0.0f, tn), b = tf;
		if (a <= b && !isinf(b) && !isinf(a)) {
			hit.tNear = a;
			hit.tFar = b;
			hit.p = ray.at(a);
			hit.n = -ray.d;
			hit.u = hit.v = 0;
			hit.prim = 0;
			tMax = a;
			did = true;
		}
	} else if (in.type == PRIM_RECTANGLE || in.type == PRIM_SPHERE || (in.type == PRIM_VOLUME && !in_lights)) {
		Hit tmp;
		bool local;
		if (in.type == PRIM_VOLUME) {
			// the proxy AABB of a HomogeneousMedia is an ordinary primitive (the dynamic_cast<GridMedia*> at Model.cpp:395 fails)
			local = aabb_intersect(V3(-0.5f, -0.5f, -0.5f), V3(0.5f, 0.5f, 0.5f), ray, tmp.tNear, tmp.tFar, tmp.n, tmp.p);
			tmp.u = tmp.v = 0;
			tmp.prim = 0;
		} else {
			local = in.type == PRIM_RECTANGLE ? rect_intersect(ray, tmp) : sphere_intersect(ray, in.radius, tmp);
		}
		if (local && !in_lights && tmp.tNear != tmp.tFar && tmp.tNear < 0 && tmp.tFar > 0) {
			// "inside" case, Model.cpp:418-424 (Q14): hit at t=0, normal left as computed at the negative root
			tmp.tNear = 0;
			tmp.p = ray.o;
			tMax = 0;
			hit = tmp;
			did = true;
		}
		if (local && tmp.tNear > tMin && tmp.tNear < tMax) {
			tMax = tmp.tNear;
			hit = tmp;
			did = true;
		}
	}
	// PRIM_POINT: Point::intersect never hits (primitives/Point.cpp:10-12)
	if (did) hit_to_wcs(in, i, hit);
	return did;
}
NE_D bool instance_intersect(const DScene& s, int i, Ray rayW, Hit& hit, float tMin, float& tMax, bool in_lights, Stats& st) {
	bool deferred;
	Ray rayO;
	return instance_intersect_t<ISECT_FULL>(s, i, rayW, hit, tMin, tMax, in_lights, st, deferred, rayO);
}

// Scene::intersectScene, core/Scene.cpp:30-56: sequential fold, instancedModels then lights, tMax shrinks on
// every accepted hit (volumes may raise it again, Q4).
NE_D bool intersect_scene(const DScene& s, Ray ray, Hit& hit, float tMin, float tMax, Stats& st) {
	bool did = false;
	hit.inst = -1;
	for (int i = 0; i < s.n_inst; i++) {
		bool th = instance_intersect(s, i, ray, hit, tMin, tMax, i >= s.n_models, st);
		did = did || th;
	}
	return did;
}
// intersect_scene for a scene the host knows to hold no triangle mesh.
NE_D bool intersect_scene_nomesh(const DScene& s, Ray ray, Hit& hit, float tMin, float tMax, Stats& st) {
	bool did = false;
	hit.inst = -1;
	for (int i = 0; i < s.n_inst; i++) {
		bool deferred;
		Ray rayO;
		bool th = instance_intersect_t<ISECT_NOMESH>(s, i, ray, hit, tMin, tMax, i >= s.n_models, st, deferred, rayO);
		did = did || th;
	}
	return did;
}

// Scene::intersectScene as a RESUMABLE computation, for the persistent trace kernels (ne_wavefront.cu, k_wf_trace): the
// fold over instances runs until a mesh needs its BVH (fold() returns true), the BVH is walked in bounded slices
// (walk()), the mesh is completed (mesh_done()) and the fold goes on. Same functions, same order of operations and
// same results as intersect_scene; what changes is that the lanes of a warp can stand at different rays.
struct SceneTrace {
	Ray rayW, rayO;
	Hit hit;
	float tMin, tMax;
	int i;  // instance the fold stands at
	bool did;
	// BVH walk of instance i
	V3 inv;
	int sp, node, slotBest;
	float tBest;
	bool any;
	// the traversal stack (int[NE_BVH_STACK], dynamically indexed, hence in local memory) is the caller's and is passed to
	// fold / walk: as a member it would drag the whole struct into local memory

	NE_D void begin(Ray r, float tMin_, float tMax_) {
		rayW = r;
		tMin = tMin_;
		tMax = tMax_;
		i = 0;
		did = false;
		hit.inst = -1;
	}
	// true: instance i is a mesh waiting for walk(); false: the fold is complete (did / hit hold the answer)
	NE_D bool fold(const DScene& s, int* stack, Stats& st) {
		while (i < s.n_inst) {
			bool deferred = false;
			bool th = instance_intersect_t<ISECT_DEFER>(s, i, rayW, hit, tMin, tMax, i >= s.n_models, st, deferred, rayO);
			if (deferred) {
				inv = bvh_inv_dir(rayO);
				sp = 0;
				stack[sp++] = NE_BVH_DONE;
				node = 0;
				tBest = INFINITY;
				slotBest = -1;
				any = false;
				return true;
			}
			did = did || th;
			i++;
		}
		return false;
	}
	// At most `budget` inner-node visits of the while-while traversal; true when the walk is complete.
	NE_D bool walk(const DMesh& m, int* stack, int budget, Stats& st) {
		while (true) {
			while (node >= 0) {
				if (budget-- <= 0) return false;
				st.bvh_nodes++;
				node = bvh_node_step(m, rayO, inv, node, stack, sp, tBest);
			}
			if (node == NE_BVH_DONE) return true;
			bvh_leaf(m, rayO, node, tBest, slotBest, any, st);
			node = stack[--sp];
		}
	}
	NE_D void mesh_done(const DScene& s) {
		const DInstance& in = s.inst[i];
		bool th = mesh_accept(s, in, s.mesh[in.mesh], rayO, any, tBest, slotBest, hit, tMin, tMax);
		if (th) hit_to_wcs(in, i, hit);
		did = did || th;
		i++;
	}
};

// ---------------------------------------------------------------------------------------------------------------
// BSDFs: BSDF wrapper core/BSDF.h:100-142; GlossyBSDF core/GlossyBSDF.cpp:10-48; GGX core/Microfacet.cpp:5-65;
// Schlick core/BSDF.h:150-157; VolumeBSDF core/VolumeBSDF.cpp:13-23; phase functions materials/Medium.h:43-129
// ---------------------------------------------------------------------------------------------------------------
NE_D float ggx_D(float alpha, V3 h) {
	float a2 = alpha * alpha;
	float NdotH = h.z;
	float NdotH2 = NdotH * NdotH;
	float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
	double dd = NE_PI * double(denom) * double(denom);
	denom = float((NE_EPSILON < dd) ? dd : NE_EPSILON);
	return a2 / denom;
}
NE_D float ggx_g1(float alpha, float NdotV) {
	float k = alpha / 2.0f;
	double dd = double(NdotV) * (1.0 - double(k)) + double(k);
	float denom = float((NE_EPSILON < dd) ? dd : NE_EPSILON);
	return NdotV / denom;
}
NE_D float ggx_G(float alpha, V3 wo, V3 wi) {
	V3 n = normalize(wo + wi);
	float NdotV = gmax(dot(n, wo), 0.0f), NdotL = gmax(dot(n, wi), 0.0f);
	return ggx_g1(alpha, NdotV) * ggx_g1(alpha, NdotL);
}
NE_D float ggx_pdf(float alpha, V3 wi, V3 h) {
	float NdotH = h.z;
	float VdotH = dot(wi, h);
	float D = ggx_D(alpha, h) * NdotH;
	return D / (4 * VdotH);
}
NE_D float fresnel_schlick(float c) { return 0.04f + (1.0f - 0.04f) * powf(1.0f - c, 5.0f); }

template <class R>
NE_D V3 ggx_sample_microfacet(float alpha, R& rng) {
	float ry = rng.next(), rx = rng.next();  // glm::vec2(random(), random()): arguments evaluated right to left
	float a2 = alpha * alpha;
	float theta = acosf(gmin(1.0f, sqrtf((1.0f - rx) / gmax(float(NE_EPSILON), (rx * (a2 - 1.0f) + 1.0f)))));
	float phi = float(NE_TWO_PI * double(ry));
	// unqualified sin/cos resolve to the double overloads in the reference build; products are formed in double
	double st, ct, sp, cp;
	sincos(double(theta), &st, &ct);
	sincos(double(phi), &sp, &cp);
	return V3(float(st * cp), float(st * sp), float(ct));
}

// ---------------------------------------------------------------------------------------------------------------
// FAST medium shading (template parameter FAST of the phase-function entry points below).
// The reference-order code divides and takes square roots in IEEE arithmetic and evaluates a phase function in a local
// frame built with generateOrthonormalCS - 57 % of k_wf_scatter's instructions on the headline frame. Production renders
// (sinks with kFast: the wavefront's volume shading and the one-thread-per-path check renderer alike) use the versions below
// instead: hardware reciprocal / reciprocal square root / sine / cosine (relative error <= 2^-21), the cosine between the
// two directions taken directly (the local frame is a rotation: it cannot change it), values that are exact by
// construction not computed at all (HG with g = 0, fr / pdf of a phase function). Only quantities that the per-function
// tests hold to 1e-5 are touched; every comparison, epsilon and ray-primitive test keeps the reference's arithmetic.
// The tape tests against the reference run FAST = false; test_fast_medium_shading_* hold FAST = true to the same oracle.
// ---------------------------------------------------------------------------------------------------------------
NE_D float rcp_fast(float x) { return __fdividef(1.0f, x); }
NE_D float sqrt_fast(float x) {
	float r;
	asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
NE_D V3 normalize_fast(V3 a) { return a * rsqrtf(dot(a, a)); }
NE_D void onb_fast(V3 n, V3& v, V3& u) {  // generateOrthonormalCS, as onb()
	if (fabsf(n.x) > fabsf(n.y))
		v = V3(-n.z, 0.0f, n.x) * rsqrtf(n.x * n.x + n.z * n.z);
	else
		v = V3(0.0f, n.z, -n.y) * rsqrtf(n.y * n.y + n.z * n.z);
	u = normalize_fast(cross(n, v));
}
// Phase function value for the pair of WORLD directions bsdf_eval / bsdf_pdf are given (wo = -incoming, wi = scattered).
NE_D float phase_eval_fast(const DMaterial& m, V3 incoming, V3 scattered) {
	if (m.phase != 1) return float(1.0 / NE_FOUR_PI);
	const float g = m.g;
	if (g == 0.0f) return NE_INV4PI;  // what hg_eval yields bit for bit: denom = 1
	float c = -dot(incoming, scattered) * rsqrtf(dot(incoming, incoming) * dot(scattered, scattered));
	float denom = 1.0f + g * g - 2.0f * g * c;
	return NE_INV4PI * (1.0f - g * g) * rcp_fast(denom) * rsqrtf(denom);
}
template <class R>
NE_D V3 phase_sample_fast(const DMaterial& m, R& rng) {
	float cosT, turn;  // turn: the azimuth in [0, 1)
	if (m.phase == 1) {
		float u1 = rng.next(), u0 = rng.next();
		float g = m.g;
		if (fabsf(g) < 1e-3f)
			cosT = 1.0f - 2.0f * u0;
		else {
			float sqr = (1.0f - g * g) * rcp_fast(1.0f + g - 2.0f * g * u0);
			sqr = sqr * sqr;
			cosT = -rcp_fast(2.0f * g) * (1.0f + g * g - sqr);
		}
		turn = u1;
	} else {
		float e2 = rng.next(), e1 = rng.next();  // sampleUnitSphere(e1, e2): cos(acos(1 - 2 e2)) = 1 - 2 e2
		cosT = 1.0f - 2.0f * e2;
		turn = e1;
	}
	// (the reference takes sqrt(1 - cos^2) of a cosine that rounding can put a hair beyond 1: a NaN direction about once in
	// 1e7 draws, tests/test_gpu_render.py CASES; clamped here)
	float sinT = sqrt_fast(fmaxf(0.0f, 1.0f - cosT * cosT));
	float sn, cs;
	__sincosf(float(NE_TWO_PI) * (turn - 0.5f), &sn, &cs);  // argument in [-pi, pi): where the hardware sine is most accurate
	return V3(-cs * sinT, -sn * sinT, cosT);  // cos(phi) = -cos(phi - pi)
}

NE_D float hg_eval(float g, V3 in, V3 out) {
	float c = dot(normalize(in), normalize(out));
	float denom = 1.0f + g * g - 2.0f * g * c;
	return NE_INV4PI * (1.0f - g * g) / (denom * sqrtf(denom));
}
NE_D float phase_eval(const DMaterial& m, V3 in, V3 out) {
	if (m.phase == 1) return hg_eval(m.g, in, out);
	return float(1.0 / NE_FOUR_PI);
}
template <class R>
NE_D V3 phase_sample(const DMaterial& m, R& rng) {
	if (m.phase == 1) {
		float u1 = rng.next(), u0 = rng.next();
		float g = m.g, cosT;
		if (fabsf(g) < 1e-3f)
			cosT = 1.0f - 2.0f * u0;
		else {
			float sqr = (1.0f - g * g) / (1.0f + g - 2.0f * g * u0);
			sqr = sqr * sqr;
			cosT = -1.0f / (2.0f * g) * (1.0f + g * g - sqr);
		}
		float sinT = sqrtf(1.0f - cosT * cosT);
		float phi = float(2.0 * NE_PI * double(u1));
		return V3(cosf(phi) * sinT, sinf(phi) * sinT, cosT);
	}
	float e2 = rng.next(), e1 = rng.next();
	return sample_unit_sphere(e1, e2);
}

// KIND: -1 = look the material's kind up at run time; 0 = known surface (GGX); 1 = known medium (phase function). A
// kernel that only ever sees one kind (k_wf_scatter: media, k_wf_surface: surfaces) compiles the other half out.
// FAST (media only, KIND = 1): see "FAST medium shading" above.
template <int KIND = -1, bool FAST = false, class R>
NE_D V3 bsdf_sample(const DScene& s, const DMaterial& m, V3 incoming, V3 normal, const Hit& ri, R& rng) {
	static_assert(!FAST || KIND == 1, "FAST shading exists for media only");
	V3 ss, ts;
	if (FAST) {
		onb_fast(normal, ss, ts);
		return to_world(phase_sample_fast(m, rng), normal, ss, ts);
	}
	onb(normal, ss, ts);
	V3 wo = to_lcs(-normalize(incoming), normal, ss, ts);
	V3 sc;
	const bool transmissive = KIND < 0 ? bool(m.transmissive) : KIND == 1;
	if (transmissive)
		sc = phase_sample(m, rng);
	else {
		float rough = material_sample(s, m.roughness_tex, ri.u, ri.v).x;
		float alpha = rough * rough;
		V3 h = ggx_sample_microfacet(alpha, rng);
		sc = reflect(normalize(wo), normalize(h));  // Q17
	}
	return to_world(sc, normal, ss, ts);
}
template <int KIND = -1, bool FAST = false>
NE_D float bsdf_pdf(const DScene& s, const DMaterial& m, V3 incoming, V3 scattered, V3 normal, const Hit& ri) {
	static_assert(!FAST || KIND == 1, "FAST shading exists for media only");
	if (FAST) return phase_eval_fast(m, incoming, scattered);
	const bool transmissive = KIND < 0 ? bool(m.transmissive) : KIND == 1;
	if (!transmissive && (!(dot(-incoming, normal) > 0) || !(dot(scattered, normal) > 0))) return 0;
	V3 ss, ts;
	onb(normal, ss, ts);
	V3 wo = to_lcs(-normalize(incoming), normal, ss, ts), wi = to_lcs(scattered, normal, ss, ts);
	if (transmissive) return phase_eval(m, wo, wi);
	V3 h = normalize(wo + wi);
	float rough = material_sample(s, m.roughness_tex, ri.u, ri.v).x;
	return ggx_pdf(rough * rough, wi, h);
}
template <int KIND = -1, bool FAST = false>
NE_D V3 bsdf_eval(const DScene& s, const DMaterial& m, V3 incoming, V3 scattered, const Hit& ri) {
	static_assert(!FAST || KIND == 1, "FAST shading exists for media only");
	if (FAST) return V3(phase_eval_fast(m, incoming, scattered));
	const bool transmissive = KIND < 0 ? bool(m.transmissive) : KIND == 1;
	if (!transmissive && (!(dot(-incoming, ri.n) > 0) || !(dot(scattered, ri.n) > 0))) return V3(0.0f);
	V3 ss, ts;
	onb(ri.n, ss, ts);
	V3 wo = to_lcs(-normalize(incoming), ri.n, ss, ts), wi = to_lcs(scattered, ri.n, ss, ts);
	if (transmissive) return V3(phase_eval(m, wo, wi));
	float rough = material_sample(s, m.roughness_tex, ri.u, ri.v).x;
	float alpha = rough * rough;
	V3 H = normalize(wo + wi);
	float HdotV = dot(wo, H);
	float NdotV = wo.z, NdotL = wi.z;
	float D = ggx_D(alpha, H);
	float G = ggx_G(alpha, wo, wi);
	float F = fresnel_schlick(HdotV);
	float num = D * G * F;
	float den = 4.0f * gmax(NdotV, 0.0f) * gmax(NdotL, 0.0f);
	float spec = num / gmax(den, float(NE_EPSILON));
	float kD = 1.0f - F;
	float metallic = material_sample(s, m.metallic_tex, ri.u, ri.v).x;
	V4 al = material_sample(s, m.albedo_tex, ri.u, ri.v);
	kD *= float(1.0 - double(metallic));
	const float pi = float(NE_PI);
	return V3(((al.x / pi) * kD + spec) * NdotL, ((al.y / pi) * kD + spec) * NdotL, ((al.z / pi) * kD + spec) * NdotL);
}

// ---------------------------------------------------------------------------------------------------------------
// GridMedia: density / interpolatedDensity / fromOCStoGCS (materials/GridMedia.cpp:15-43, GridMedia.h:33-36) over
// the brick-sparse grid; Tr (ratio tracking) :45-69; sample (delta tracking) :71-100.
// ---------------------------------------------------------------------------------------------------------------
NE_D float interpolated_density(const DVolume& v, V3 g) {
	g = gmin(gmax(V3(0.0f), g), V3(float(v.W), float(v.H), float(v.D)));
	int ix = int(floorf(g.x)), iy = int(floorf(g.y)), iz = int(floorf(g.z));
	if (ix >= v.W || iy >= v.H || iz >= v.D) return 0.0f;  // every corner is at or beyond the grid: density() = 0
	int b = __ldg(&v.cells[((iz >> 3) * v.by + (iy >> 3)) * v.bx + (ix >> 3)].x);
	if (b < 0) return 0.0f;
	V3 d = g - V3(float(ix), float(iy), float(iz));
	// apron layout: all eight corners are in this brick's 9^3 record
	const float* p = v.pool + size_t(b) * BRICK_VOX + ((iz & 7) * 9 + (iy & 7)) * 9 + (ix & 7);
	float v000 = __ldg(p), v100 = __ldg(p + 1);
	float v010 = __ldg(p + 9), v110 = __ldg(p + 10);
	float v001 = __ldg(p + 81), v101 = __ldg(p + 82);
	float v011 = __ldg(p + 90), v111 = __ldg(p + 91);
	float d00 = gmix(v000, v100, d.x), d10 = gmix(v010, v110, d.x), d01 = gmix(v001, v101, d.x), d11 = gmix(v011, v111, d.x);
	float d0 = gmix(d00, d10, d.y), d1 = gmix(d01, d11, d.y);
	return gmix(d0, d1, d.z);
}
NE_D V3 ocs_to_gcs(const DVolume& v, V3 p) { return (p + V3(0.5f)) * V3(float(v.W), float(v.H), float(v.D)); }
NE_D float density_at(const DVolume& v, const Ray& rayO, float t) { return interpolated_density(v, ocs_to_gcs(v, rayO.at(t))); }

// Trilinear density at grid point g, known to lie in brick (bx,by,bz) whose record is `slot` (>= 0): no table
// look-up, eight loads from ONE 9^3 record. The local cell is clamped into the record, so a point that rounding put a
// hair outside the brick is evaluated on the brick's own face (the brick majorant still bounds it).
NE_D float brick_density(const float* __restrict__ pool, int slot, V3 g, int bx, int by, int bz) {
	float lx = g.x - float(bx << 3), ly = g.y - float(by << 3), lz = g.z - float(bz << 3);
	lx = fminf(fmaxf(lx, 0.0f), 8.0f); ly = fminf(fmaxf(ly, 0.0f), 8.0f); lz = fminf(fmaxf(lz, 0.0f), 8.0f);
	int ix = min(int(lx), 7), iy = min(int(ly), 7), iz = min(int(lz), 7);
	float fx = lx - float(ix), fy = ly - float(iy), fz = lz - float(iz);
	const float* p = pool + size_t(slot) * BRICK_VOX + (iz * 9 + iy) * 9 + ix;
	float v000 = __ldg(p), v100 = __ldg(p + 1);
	float v010 = __ldg(p + 9), v110 = __ldg(p + 10);
	float v001 = __ldg(p + 81), v101 = __ldg(p + 82);
	float v011 = __ldg(p + 90), v111 = __ldg(p + 91);
	float d00 = fmaf(fx, v100 - v000, v000), d10 = fmaf(fx, v110 - v010, v010);
	float d01 = fmaf(fx, v101 - v001, v001), d11 = fmaf(fx, v111 - v011, v011);
	float d0 = fmaf(fy, d10 - d00, d00), d1 = fmaf(fy, d11 - d01, d01);
	return fmaf(fz, d1 - d0, d0);
}

}  // namespace ne
#include "ne_tracking.cuh"  // Tracker, ratio_walk, delta_walk, grid_tr, grid_scatter, grid_sample
namespace ne {


// ---------------------------------------------------------------------------------------------------------------
// Emitter primitives: samplePointOnSurface / pdf of Rectangle (Rectangle.cpp:78-88,159-171), Sphere
// (Sphere.cpp:46-52,58-68), Point (Point.cpp:14-26). `prim` = the primitive Light::primitive points at (Q7),
// `li` = the chosen light instance (its transform and scale are used).
// ---------------------------------------------------------------------------------------------------------------
// LS: -1 = look the emitter's primitive type up at run time; PRIM_RECTANGLE / PRIM_SPHERE / PRIM_POINT = the host found
// that EVERY light of the scene is a DiffuseLight on that kind of primitive (ctx->lightSet), so the shading kernels compile
// the other kinds - and the directional / environment light code - out (k_wf_scatter was 17.6 K instructions, 280 KB of
// code: instruction-cache misses were among its top stalls).
template <int LS = -1, class R>
NE_D V3 light_sample_point(const DInstance& prim, const DInstance& li, const Hit& isect, R& rng) {
	const int type = LS >= 0 ? LS : prim.type;
	if (type == PRIM_RECTANGLE) {
		float e2 = rng.next(), e1 = rng.next(), e0 = rng.next();  // vec3(random(),random(),random()) right to left
		V3 p(-0.5f + 1.0f * e0, -0.5f + 1.0f * e1, 0.0f + 0.0f * e2);
		return xform_point(li.M, p);
	}
	if (type == PRIM_SPHERE) {
		V3 c = xform_point(li.M, V3(0.0f, 0.0f, 0.0f));
		V3 n = normalize(isect.p - c);
		return c + n * prim.radius;
	}
	return xform_point(li.M, V3(prim.point[0], prim.point[1], prim.point[2]));
}
template <int LS = -1, class R>
NE_D float light_pdf(const DInstance& prim, const DInstance& li, const Hit& isect, R& rng) {
	const int type = LS >= 0 ? LS : prim.type;
	if (type == PRIM_RECTANGLE) {
		V3 sizeW = V3(li.scale[0], li.scale[1], li.scale[2]) * V3(1.0f, 1.0f, 0.0f);
		float area = 1;
		if (sizeW.x != 0) area = area * sizeW.x;
		if (sizeW.y != 0) area = area * sizeW.y;
		if (sizeW.z != 0) area = area * sizeW.z;
		return area_to_solid_angle(1.0f / area, isect.n, isect.p, light_sample_point<LS>(prim, li, isect, rng));  // Q10, Q11
	}
	if (type == PRIM_SPHERE) {
		float pdfArea = float(double(1.0f / 4.0f) * NE_PI * double(prim.radius) * double(prim.radius));  // Q13
		return area_to_solid_angle(pdfArea, isect.n, isect.p, light_sample_point<LS>(prim, li, isect, rng));
	}
	return 1;
}

// Camera::getRayPassingThrough, core/Camera.cpp:140-144 + randomInUnitDisk, utils/Math.h:502-507
template <class R>
NE_D Ray camera_ray(const DCamera& c, float x, float y, R& rng) {
	float theta = float(2.0 * NE_PI * double(rng.next()));
	float r = sqrtf(rng.next());
	double sn, cs;
	sincos(double(theta), &sn, &cs);  // ::cos/::sin(double) in the reference build
	V3 rd = c.lens_radius * (r * V3(float(cs), float(sn), 0.0f));
	V3 offset = c.side * rd.x + c.up * rd.y;
	Ray ray;
	ray.o = c.position + offset;
	ray.d = -normalize(c.lower_left + x * c.horizontal + y * c.vertical - c.position - offset);
	return ray;
}

// OfflineEngine::postProcessing, core/OfflineEngine.cpp:39-52 (exposure 0.5, gamma 2.2, clamp)
NE_D float tonemap1(float c) {
	float m = 1.0f - expf(-c * 0.5f);
	m = powf(m, 1.0f / 2.2f);
	return gclamp(m, 0.0f, 1.0f);
}

}  // namespace ne
