// ne_host.cpp — host-side scene preparation. See ne_host.h.
#include "ne_host.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>

namespace ne {
void set_error(const std::string& s);  // ne_api.cu

// ---------------------------------------------------------------------------------------------------------------
// Transform arithmetic in glm 0.9.9.4's operation order (includes/glm):
//   getTransform = translate(I, T) * eulerAngleXYZ(radians(R)) then scale(., S)   src/utils/Math.h:848-859
//   glm::inverse (cofactor form)                                                 glm/detail/func_matrix.inl:294-352
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Mat {
	float c[4][4];  // c[col][row]
};
Mat identity() {
	Mat m;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) m.c[i][j] = i == j ? 1.0f : 0.0f;
	return m;
}
Mat mul(const Mat& a, const Mat& b) {
	Mat r;
	for (int j = 0; j < 4; j++)
		for (int k = 0; k < 4; k++) r.c[j][k] = ((a.c[0][k] * b.c[j][0] + a.c[1][k] * b.c[j][1]) + a.c[2][k] * b.c[j][2]) + a.c[3][k] * b.c[j][3];
	return r;
}
Mat inverse(const Mat& mm) {
	const float(*m)[4] = mm.c;
	float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07}, Fac2[4] = {Coef08, Coef08, Coef10, Coef11};
	float Fac3[4] = {Coef12, Coef12, Coef14, Coef15}, Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
	float Vec0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, Vec1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
	float Vec2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, Vec3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
	const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
	Mat inv;
	for (int i = 0; i < 4; i++) {
		float Inv0 = (Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i]) + Vec3[i] * Fac2[i];
		float Inv1 = (Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i]) + Vec3[i] * Fac4[i];
		float Inv2 = (Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i]) + Vec3[i] * Fac5[i];
		float Inv3 = (Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i]) + Vec2[i] * Fac5[i];
		inv.c[0][i] = Inv0 * SignA[i];
		inv.c[1][i] = Inv1 * SignB[i];
		inv.c[2][i] = Inv2 * SignA[i];
		inv.c[3][i] = Inv3 * SignB[i];
	}
	float Dot0[4] = {m[0][0] * inv.c[0][0], m[0][1] * inv.c[1][0], m[0][2] * inv.c[2][0], m[0][3] * inv.c[3][0]};
	float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
	float ood = 1.0f / Dot1;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) inv.c[i][j] = inv.c[i][j] * ood;
	return inv;
}
}  // namespace

void host_make_transform(const float pos[3], const float rotDeg[3], const float scale[3], float M[16], float Mi[16]) {
	Mat model = identity();
	// glm::translate: Result[3] = m[0]*v[0] + m[1]*v[1] + m[2]*v[2] + m[3]
	for (int k = 0; k < 4; k++) model.c[3][k] = ((model.c[0][k] * pos[0] + model.c[1][k] * pos[1]) + model.c[2][k] * pos[2]) + model.c[3][k];
	// glm::eulerAngleXYZ(radians)
	const float rad = static_cast<float>(0.01745329251994329576923690768489);
	float t1 = rotDeg[0] * rad, t2 = rotDeg[1] * rad, t3 = rotDeg[2] * rad;
	float c1 = std::cos(-t1), c2 = std::cos(-t2), c3 = std::cos(-t3), s1 = std::sin(-t1), s2 = std::sin(-t2), s3 = std::sin(-t3);
	Mat R = identity();
	R.c[0][0] = c2 * c3;
	R.c[0][1] = -c1 * s3 + s1 * s2 * c3;
	R.c[0][2] = s1 * s3 + c1 * s2 * c3;
	R.c[1][0] = c2 * s3;
	R.c[1][1] = c1 * c3 + s1 * s2 * s3;
	R.c[1][2] = -s1 * c3 + c1 * s2 * s3;
	R.c[2][0] = -s2;
	R.c[2][1] = s1 * c2;
	R.c[2][2] = c1 * c2;
	model = mul(model, R);
	for (int k = 0; k < 4; k++) {
		model.c[0][k] = model.c[0][k] * scale[0];
		model.c[1][k] = model.c[1][k] * scale[1];
		model.c[2][k] = model.c[2][k] * scale[2];
	}
	Mat inv = inverse(model);
	memcpy(M, model.c, 64);
	memcpy(Mi, inv.c, 64);
}

// getScale, src/utils/Math.h:874-912 (glm::decompose's scale extraction)
void host_get_scale(const float M[16], float s[3]) {
	float l[4][4];
	memcpy(l, M, 64);
	float w = l[3][3];
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) l[i][j] /= w;  // the divisor is read through the reference too; [3][3] is divided last
	float row[3][3];
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) row[i][j] = l[i][j];
	auto len = [](const float* v) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); };
	auto dot3 = [](const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; };
	auto scaleTo = [&](float* v, float desired) {  // glm::detail::scale: v * desired / length(v)
		float L = len(v);
		for (int k = 0; k < 3; k++) v[k] = v[k] * desired / L;
	};
	auto combine = [](float* a, const float* b, float as, float bs) {  // (a * as) + (b * bs)
		for (int k = 0; k < 3; k++) a[k] = (a[k] * as) + (b[k] * bs);
	};
	s[0] = len(row[0]);
	scaleTo(row[0], 1.0f);
	float skewZ = dot3(row[0], row[1]);
	combine(row[1], row[0], 1.0f, -skewZ);
	s[1] = len(row[1]);
	scaleTo(row[1], 1.0f);
	float skewY = dot3(row[0], row[2]);
	combine(row[2], row[0], 1.0f, -skewY);
	float skewX = dot3(row[1], row[2]);
	combine(row[2], row[1], 1.0f, -skewX);
	s[2] = len(row[2]);
}

// Camera::Camera, src/core/Camera.cpp:7-26
void host_camera_make(const float from[3], const float at[3], const float upv[3], float vfov, float aspect, float aperture, float focus,
                      ne_b200_camera* out) {
	V3 lookFrom(from[0], from[1], from[2]), lookAt(at[0], at[1], at[2]), up(upv[0], upv[1], upv[2]);
	float lensRadius = aperture / 2.0f;
	const float rad = static_cast<float>(0.01745329251994329576923690768489);
	float halfHeight = float(::tan(double((vfov * rad) / 2.0f)));  // ::tan(double) in the reference build
	float halfWidth = aspect * halfHeight;
	V3 position = lookFrom;
	V3 front = normalize(lookFrom - lookAt);
	V3 side = normalize(cross(up, front));
	V3 lowerLeft = position - focus * halfWidth * -side - focus * halfHeight * up - focus * -front;
	V3 horizontal = 2.0f * focus * halfWidth * -side;
	V3 vertical = 2.0f * focus * halfHeight * up;
	auto put = [](float* o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
	put(out->position, position);
	put(out->lower_left, lowerLeft);
	put(out->horizontal, horizontal);
	put(out->vertical, vertical);
	put(out->side, side);
	put(out->up, up);
	out->lens_radius = lensRadius;
}

// ---------------------------------------------------------------------------------------------------------------
// Binned-SAH BVH
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Box {
	float lo[3], hi[3];
	void reset() { lo[0] = lo[1] = lo[2] = INFINITY; hi[0] = hi[1] = hi[2] = -INFINITY; }
	void grow(const Box& b) {
		for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
	}
	void grow(const float* p) {
		for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
	}
	float area() const {
		float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
		if (dx < 0) return 0;
		return 2.0f * (dx * dy + dy * dz + dz * dx);
	}
};
struct Builder {
	const float* pos;
	const uint32_t* idx;
	std::vector<Box> tbox;
	std::vector<float> cent;  // 3 per tri
	std::vector<uint32_t> order;
	std::vector<BvhNode>* nodes;
	std::vector<uint32_t> leafOrder;  // triangle slots in leaf order
	static constexpr int NB = 16, LEAF = 4;

	// returns child reference (node index or ~((first<<3)|cnt)) and its box
	int build(uint32_t s, uint32_t e, Box& box) {
		box.reset();
		Box cb;
		cb.reset();
		for (uint32_t i = s; i < e; i++) {
			box.grow(tbox[order[i]]);
			cb.grow(&cent[3 * size_t(order[i])]);
		}
		uint32_t n = e - s;
		if (n <= LEAF) return makeLeaf(s, e);
		int axis = 0;
		float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
		if (ext[1] > ext[axis]) axis = 1;
		if (ext[2] > ext[axis]) axis = 2;
		uint32_t mid = s + n / 2;
		if (ext[axis] > 0) {
			Box bb[NB];
			uint32_t bc[NB];
			for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
			float k = NB * (1.0f - 1e-6f) / ext[axis];
			for (uint32_t i = s; i < e; i++) {
				int b = std::min(NB - 1, std::max(0, int((cent[3 * size_t(order[i]) + axis] - cb.lo[axis]) * k)));
				bb[b].grow(tbox[order[i]]);
				bc[b]++;
			}
			float rightA[NB];
			uint32_t rightC[NB];
			Box acc;
			acc.reset();
			uint32_t c = 0;
			for (int b = NB - 1; b > 0; b--) {
				acc.grow(bb[b]);
				c += bc[b];
				rightA[b] = acc.area();
				rightC[b] = c;
			}
			acc.reset();
			c = 0;
			float best = INFINITY;
			int bestB = -1;
			for (int b = 0; b < NB - 1; b++) {
				acc.grow(bb[b]);
				c += bc[b];
				if (c == 0 || rightC[b + 1] == 0) continue;
				float cost = acc.area() * c + rightA[b + 1] * rightC[b + 1];
				if (cost < best) { best = cost; bestB = b; }
			}
			if (bestB >= 0) {
				auto it = std::partition(order.begin() + s, order.begin() + e, [&](uint32_t t) {
					int b = std::min(NB - 1, std::max(0, int((cent[3 * size_t(t) + axis] - cb.lo[axis]) * k)));
					return b <= bestB;
				});
				mid = uint32_t(it - order.begin());
			}
		}
		if (mid == s || mid == e) {
			mid = s + n / 2;
			std::nth_element(order.begin() + s, order.begin() + mid, order.begin() + e,
			                 [&](uint32_t a, uint32_t b) { return cent[3 * size_t(a) + axis] < cent[3 * size_t(b) + axis]; });
		}
		int me = int(nodes->size());
		nodes->push_back(BvhNode());
		Box b0, b1;
		int c0 = build(s, mid, b0);
		int c1 = build(mid, e, b1);
		BvhNode& nd = (*nodes)[me];
		for (int k2 = 0; k2 < 3; k2++) { nd.lo0[k2] = b0.lo[k2]; nd.hi0[k2] = b0.hi[k2]; nd.lo1[k2] = b1.lo[k2]; nd.hi1[k2] = b1.hi[k2]; }
		nd.child0 = c0;
		nd.child1 = c1;
		nd.cnt0 = nd.cnt1 = 0;
		return me;
	}
	int makeLeaf(uint32_t s, uint32_t e) {
		uint32_t first = uint32_t(leafOrder.size());
		for (uint32_t i = s; i < e; i++) leafOrder.push_back(order[i]);
		return ~int((first << 3) | (e - s));
	}
};
}  // namespace

void host_build_bvh(const float* positions, int nVerts, const uint32_t* indices, int nTris, HostBvh& out) {
	Builder b;
	b.pos = positions;
	b.idx = indices;
	b.tbox.resize(nTris);
	b.cent.resize(3 * size_t(nTris));
	b.order.resize(nTris);
	b.nodes = &out.nodes;
	out.nodes.clear();
	out.nodes.reserve(size_t(nTris) / 2 + 16);
	for (int t = 0; t < nTris; t++) {
		b.order[t] = t;
		Box& bx = b.tbox[t];
		bx.reset();
		for (int k = 0; k < 3; k++) bx.grow(positions + 3 * size_t(indices[3 * size_t(t) + k]));
		for (int k = 0; k < 3; k++) b.cent[3 * size_t(t) + k] = 0.5f * (bx.lo[k] + bx.hi[k]);
	}
	// Model::boundingBox = min/max over all VERTICES (Model.cpp:160-163)
	Box all;
	all.reset();
	for (int v = 0; v < nVerts; v++) all.grow(positions + 3 * size_t(v));
	for (int k = 0; k < 3; k++) { out.bbmin[k] = all.lo[k]; out.bbmax[k] = all.hi[k]; }

	Box rootBox;
	if (nTris <= Builder::LEAF) {
		out.nodes.push_back(BvhNode());
		int leaf = nTris > 0 ? b.build(0, nTris, rootBox) : ~0;
		if (nTris == 0) rootBox.reset();
		BvhNode& nd = out.nodes[0];
		for (int k = 0; k < 3; k++) { nd.lo0[k] = rootBox.lo[k]; nd.hi0[k] = rootBox.hi[k]; nd.lo1[k] = INFINITY; nd.hi1[k] = -INFINITY; }
		nd.child0 = leaf;
		nd.child1 = ~0;
		nd.cnt0 = nd.cnt1 = 0;
	} else
		b.build(0, nTris, rootBox);

	out.tri.resize(12 * b.leafOrder.size());
	for (size_t sl = 0; sl < b.leafOrder.size(); sl++) {
		uint32_t t = b.leafOrder[sl];
		float* o = &out.tri[12 * sl];
		for (int k = 0; k < 3; k++) {
			const float* p = positions + 3 * size_t(indices[3 * size_t(t) + k]);
			o[4 * k] = p[0]; o[4 * k + 1] = p[1]; o[4 * k + 2] = p[2]; o[4 * k + 3] = 0.0f;
		}
		memcpy(&o[3], &t, 4);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Brick-sparse density grid
// ---------------------------------------------------------------------------------------------------------------
namespace {
void atomicMaxF(std::atomic<uint32_t>& a, float v) {
	if (!(v > 0)) return;
	uint32_t bits;
	memcpy(&bits, &v, 4);
	uint32_t cur = a.load(std::memory_order_relaxed);
	while (cur < bits && !a.compare_exchange_weak(cur, bits, std::memory_order_relaxed)) {}
}
template <class F>
void parallelFor(size_t n, F f) {
	unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
	if (n < 4096) nt = 1;
	std::vector<std::thread> th;
	std::atomic<size_t> next{0};
	const size_t chunk = 256;
	auto work = [&]() {
		while (true) {
			size_t s = next.fetch_add(chunk);
			if (s >= n) break;
			size_t e = std::min(n, s + chunk);
			for (size_t i = s; i < e; i++) f(i);
		}
	};
	for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
	work();
	for (auto& t : th) t.join();
}
}  // namespace

void host_build_bricks(const ne_b200_volume& v, HostBricks& out) {
	out.W = v.width; out.H = v.height; out.D = v.depth;
	out.bx = (v.width + 7) / 8; out.by = (v.height + 7) / 8; out.bz = (v.depth + 7) / 8;
	size_t nb = size_t(out.bx) * out.by * out.bz;
	out.table.assign(nb, -1);
	out.pool.clear();
	const size_t W = v.width, H = v.height;

	if (v.dense) {
		// pass 1: which bricks hold a non-zero voxel
		std::vector<uint8_t> active(nb, 0);
		parallelFor(nb, [&](size_t b) {
			int bx = int(b % out.bx), by = int((b / out.bx) % out.by), bz = int(b / (size_t(out.bx) * out.by));
			for (int z = bz * 8; z < std::min(v.depth, bz * 8 + 8); z++)
				for (int y = by * 8; y < std::min(v.height, by * 8 + 8); y++) {
					const float* row = v.dense + W * H * z + W * y;
					for (int x = bx * 8; x < std::min(v.width, bx * 8 + 8); x++)
						if (row[x] != 0.0f) { active[b] = 1; return; }
				}
		});
		int32_t slots = 0;
		for (size_t b = 0; b < nb; b++)
			if (active[b]) out.table[b] = slots++;
		out.pool.assign(size_t(slots) * CORE_VOX, 0.0f);
		parallelFor(nb, [&](size_t b) {
			if (out.table[b] < 0) return;
			int bx = int(b % out.bx), by = int((b / out.bx) % out.by), bz = int(b / (size_t(out.bx) * out.by));
			float* dst = &out.pool[size_t(out.table[b]) * CORE_VOX];
			for (int z = 0; z < 8 && bz * 8 + z < v.depth; z++)
				for (int y = 0; y < 8 && by * 8 + y < v.height; y++) {
					const float* row = v.dense + W * H * (bz * 8 + z) + W * (by * 8 + y);
					for (int x = 0; x < 8 && bx * 8 + x < v.width; x++) dst[64 * z + 8 * y + x] = row[bx * 8 + x];
				}
		});
	} else {
		int32_t slots = 0;
		std::vector<int32_t> leafSlot(std::max(0, v.n_leaves), -1);
		for (int l = 0; l < v.n_leaves; l++) {
			const int32_t* o = v.leaf_origin + 3 * size_t(l);
			if (o[0] < 0 || o[1] < 0 || o[2] < 0 || o[0] >= v.width || o[1] >= v.height || o[2] >= v.depth) continue;
			if ((o[0] | o[1] | o[2]) & 7) continue;  // leaves are 8-aligned (documented in ne_b200.h)
			size_t b = (size_t(o[2] / 8) * out.by + o[1] / 8) * out.bx + o[0] / 8;
			if (out.table[b] < 0) out.table[b] = slots++;
			leafSlot[l] = out.table[b];
		}
		out.pool.assign(size_t(slots) * CORE_VOX, 0.0f);
		// later leaves overwrite earlier ones at the same origin, like repeated copyToDense writes would
		for (int l = 0; l < v.n_leaves; l++) {
			if (leafSlot[l] < 0) continue;
			const int32_t* o = v.leaf_origin + 3 * size_t(l);
			const float* src = v.leaf_values + size_t(CORE_VOX) * l;
			float* dst = &out.pool[size_t(leafSlot[l]) * CORE_VOX];
			for (int z = 0; z < 8; z++)
				for (int y = 0; y < 8; y++)
					for (int x = 0; x < 8; x++) {
						bool in = o[0] + x < v.width && o[1] + y < v.height && o[2] + z < v.depth;
						dst[64 * z + 8 * y + x] = in ? src[64 * z + 8 * y + x] : 0.0f;
					}
		}
	}

	// Global maximum (GridMedia::calculateMaxDensity starts from 0, GridMedia.h:16-21) and per-brick majorants over
	// the [8b-1, 8b+8] support of each brick (the trilinear stencil of any point inside the brick, +-1 voxel).
	std::vector<std::atomic<uint32_t>> maj(nb);
	for (auto& a : maj) a.store(0, std::memory_order_relaxed);
	std::atomic<uint32_t> gmaxBits{0};
	parallelFor(nb, [&](size_t b) {
		int32_t slot = out.table[b];
		if (slot < 0) return;
		int bx = int(b % out.bx), by = int((b / out.bx) % out.by), bz = int(b / (size_t(out.bx) * out.by));
		const float* src = &out.pool[size_t(slot) * CORE_VOX];
		// range of local indices contributing to the neighbour at offset -1 / 0 / +1 along one axis
		const int lo[3] = {0, 0, 7}, hi[3] = {0, 7, 7};
		for (int oz = -1; oz <= 1; oz++)
			for (int oy = -1; oy <= 1; oy++)
				for (int ox = -1; ox <= 1; ox++) {
					int nx = bx + ox, ny = by + oy, nz = bz + oz;
					if (nx < 0 || ny < 0 || nz < 0 || nx >= out.bx || ny >= out.by || nz >= out.bz) continue;
					float m = 0;
					for (int z = lo[oz + 1]; z <= hi[oz + 1]; z++)
						for (int y = lo[oy + 1]; y <= hi[oy + 1]; y++)
							for (int x = lo[ox + 1]; x <= hi[ox + 1]; x++) m = std::max(m, src[64 * z + 8 * y + x]);
					atomicMaxF(maj[(size_t(nz) * out.by + ny) * out.bx + nx], m);
					if (ox == 0 && oy == 0 && oz == 0) atomicMaxF(gmaxBits, m);
				}
	});
	out.bmaj.resize(nb);
	for (size_t b = 0; b < nb; b++) {
		uint32_t bits = maj[b].load(std::memory_order_relaxed);
		memcpy(&out.bmaj[b], &bits, 4);
	}
	uint32_t gb = gmaxBits.load();
	memcpy(&out.maxDensity, &gb, 4);
	// 1/majorant (0 = nothing to collide with in this brick): the tracking loop never divides
	out.binv.resize(nb);
	for (size_t b = 0; b < nb; b++) out.binv[b] = out.bmaj[b] > 0 ? 1.0f / out.bmaj[b] : 0.0f;

	// Apron layout: every stored brick holds the 9x9x9 voxels [8b, 8b+8]^3, so the eight corners of any cell of
	// the brick come from ONE brick record (one table look-up, no neighbour fetches). A brick gets storage iff that
	// 9^3 support holds a non-zero voxel. Voxels at or beyond the grid's size are 0 (GridMedia::density :17-18).
	std::vector<int32_t> coreTable;
	coreTable.swap(out.table);
	std::vector<float> core;
	core.swap(out.pool);
	auto coreVoxel = [&](int x, int y, int z) -> float {
		if (x >= out.W || y >= out.H || z >= out.D) return 0.0f;
		int32_t s = coreTable[(size_t(z >> 3) * out.by + (y >> 3)) * out.bx + (x >> 3)];
		return s < 0 ? 0.0f : core[size_t(s) * CORE_VOX + ((z & 7) << 6) + ((y & 7) << 3) + (x & 7)];
	};
	std::vector<uint8_t> need(nb, 0);
	parallelFor(nb, [&](size_t b) {
		int bx = int(b % out.bx), by = int((b / out.bx) % out.by), bz = int(b / (size_t(out.bx) * out.by));
		if (coreTable[b] >= 0) { need[b] = 1; return; }
		// an empty core still needs storage when a +x/+y/+z neighbour's first voxel layer is non-zero
		for (int z = 0; z <= 8; z++)
			for (int y = 0; y <= 8; y++)
				for (int x = 0; x <= 8; x++) {
					if (x < 8 && y < 8 && z < 8) continue;
					if (coreVoxel(bx * 8 + x, by * 8 + y, bz * 8 + z) != 0.0f) { need[b] = 1; return; }
				}
	});
	out.table.assign(nb, -1);
	int32_t slots = 0;
	for (size_t b = 0; b < nb; b++)
		if (need[b]) out.table[b] = slots++;
	out.pool.assign(size_t(slots) * BRICK_VOX, 0.0f);
	parallelFor(nb, [&](size_t b) {
		if (out.table[b] < 0) return;
		int bx = int(b % out.bx), by = int((b / out.bx) % out.by), bz = int(b / (size_t(out.bx) * out.by));
		float* dst = &out.pool[size_t(out.table[b]) * BRICK_VOX];
		for (int z = 0; z <= 8; z++)
			for (int y = 0; y <= 8; y++)
				for (int x = 0; x <= 8; x++) dst[(z * 9 + y) * 9 + x] = coreVoxel(bx * 8 + x, by * 8 + y, bz * 8 + z);
	});
}

}  // namespace ne

// ---------------------------------------------------------------------------------------------------------------
// C ABI of the builders (include/ne_b200.h "Host-side scene builders")
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int ne_b200_host_build_bricks(const ne_b200_volume* v, int32_t dims[4], int32_t* table, float* inv_majorant, float* pool, float* max_density) {
	if (!v || !dims) { ne::set_error("null argument"); return NE_B200_ERR_INVALID; }
	if (v->width <= 0 || v->height <= 0 || v->depth <= 0 || (!v->dense && v->n_leaves > 0 && (!v->leaf_origin || !v->leaf_values))) {
		ne::set_error("bad volume");
		return NE_B200_ERR_INVALID;
	}
	ne::HostBricks hb;
	ne::host_build_bricks(*v, hb);
	dims[0] = hb.bx; dims[1] = hb.by; dims[2] = hb.bz;
	dims[3] = int32_t(hb.pool.size() / ne::BRICK_VOX);
	if (table) memcpy(table, hb.table.data(), hb.table.size() * sizeof(int32_t));
	if (inv_majorant) memcpy(inv_majorant, hb.binv.data(), hb.binv.size() * sizeof(float));
	if (pool) memcpy(pool, hb.pool.data(), hb.pool.size() * sizeof(float));
	if (max_density) *max_density = hb.maxDensity;
	return NE_B200_OK;
}

int ne_b200_host_build_bvh(const float* positions, int32_t n_vertices, const uint32_t* indices, int32_t n_triangles, int32_t counts[2], void* nodes,
                           float* triangles) {
	if (!counts || n_vertices < 0 || n_triangles < 0 || (n_triangles > 0 && (!positions || !indices))) { ne::set_error("bad argument"); return NE_B200_ERR_INVALID; }
	for (int t = 0; t < 3 * n_triangles; t++)
		if (indices[t] >= uint32_t(n_vertices)) { ne::set_error("mesh index out of range"); return NE_B200_ERR_INVALID; }
	ne::HostBvh hb;
	ne::host_build_bvh(positions, n_vertices, indices, n_triangles, hb);
	counts[0] = int32_t(hb.nodes.size());
	counts[1] = int32_t(hb.tri.size() / 12);
	if (nodes) memcpy(nodes, hb.nodes.data(), hb.nodes.size() * sizeof(ne::BvhNode));
	if (triangles) memcpy(triangles, hb.tri.data(), hb.tri.size() * sizeof(float));
	return NE_B200_OK;
}

}  // extern "C"
