// ne_math.cuh — fp32 vector arithmetic with the SAME operation order as the glm 0.9.9.4 scalar paths the reference
// uses (includes/glm, no GLM_FORCE_INTRINSICS), so that hit/no-hit decisions round the same way
// (SURVEY.md §7 hard part 2). The library is compiled with -fmad=false: no a*b+c contraction anywhere.
//
// Reference arithmetic restated here:
//   glm::dot(vec3)            (a.x*b.x + a.y*b.y) + a.z*b.z                       glm/detail/func_geometric.inl
//   glm::normalize            v * (1 / sqrt(dot(v,v)))                             glm/detail/func_geometric.inl
//   glm::cross, reflect       glm/detail/func_geometric.inl
//   mat4 * vec4               (m0*x + m1*y) + (m2*z + m3*w)                        glm/detail/type_mat4x4.inl:536-580
//   glm::min/max              (b<a)?b:a / (a<b)?b:a  (NaN behaviour differs from fminf/fmaxf)
//   glm::mix / lerp           x*(1-a) + y*a                                        glm/detail/func_common.inl:104-112
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define NE_HD __host__ __device__ __forceinline__
#define NE_D __device__ __forceinline__

namespace ne {

// Constants of src/utils/Math.h:18-29 (doubles unless cast).
#define NE_EPSILON3 0.001
#define NE_EPSILON 0.0000000001
#define NE_EPSILON12 0.00000000001
#define NE_PI 3.14159265358979323846264338327950288
#define NE_TWO_PI 6.283185307179586476925286766559
#define NE_FOUR_PI 12.566370614359172953850573533118
#define NE_INV4PI float(0.07957747154594766788)

struct V3 {
	float x, y, z;
	NE_HD V3() {}
	NE_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
	NE_HD explicit V3(float a) : x(a), y(a), z(a) {}
	NE_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct V4 {
	float x, y, z, w;
};

NE_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
NE_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
NE_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
NE_HD V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
NE_HD V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
NE_HD V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
NE_HD V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
NE_HD V3 operator/(float s, V3 a) { return V3(s / a.x, s / a.y, s / a.z); }
NE_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
NE_HD V3 operator+(V3 a, float s) { return V3(a.x + s, a.y + s, a.z + s); }
NE_HD V3 operator-(V3 a, float s) { return V3(a.x - s, a.y - s, a.z - s); }

NE_HD float gmin(float a, float b) { return (b < a) ? b : a; }
NE_HD float gmax(float a, float b) { return (a < b) ? b : a; }
NE_HD V3 gmin(V3 a, V3 b) { return V3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
NE_HD V3 gmax(V3 a, V3 b) { return V3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
NE_HD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
NE_HD float gsign(float x) { return float(int(0.0f < x) - int(x < 0.0f)); }
NE_HD float gstep(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
NE_HD float gmix(float x, float y, float a) { return x * (1.0f - a) + y * a; }

NE_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NE_HD V3 cross(V3 x, V3 y) { return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
NE_HD float length(V3 a) { return sqrtf(dot(a, a)); }
NE_HD float length2(V3 a) { return dot(a, a); }
NE_HD V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
NE_HD V3 reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }
NE_HD V3 vabs(V3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
NE_HD bool is_black(V3 v) { return v.x == 0 && v.y == 0 && v.z == 0; }      // Math.h:641-647
NE_HD bool all_one(V3 v) { return v.x == 1 && v.y == 1 && v.z == 1; }       // Math.h:649-651
NE_HD float avg(V3 v) { return (v.x + v.y + v.z) / 3.0f; }                  // Math.h:930-932

// Column-major 4x4 like glm::mat4: m[4*c + r].
struct M4 {
	float m[16];
};
// (M * vec4(v, 1)).xyz
NE_HD V3 xform_point(const float* m, V3 v) {
	return V3((m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * 1.0f), (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * 1.0f),
	          (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * 1.0f));
}
// (M * vec4(v, 0)).xyz
NE_HD V3 xform_dir(const float* m, V3 v) {
	return V3((m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * 0.0f), (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * 0.0f),
	          (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * 0.0f));
}

struct Ray {
	V3 o, d;
	NE_HD V3 at(float t) const { return o + t * d; }  // Ray::getPointAt, src/primitives/Ray.h:25-27
};
// transformRay, src/utils/Math.h:942-948 (direction NOT renormalised: t is shared between WCS and OCS)
NE_HD Ray transform_ray(Ray r, const float* m) {
	Ray o;
	o.o = xform_point(m, r.o);
	o.d = xform_dir(m, r.d);
	return o;
}

// generateOrthonormalCS, src/utils/Math.h:591-599
NE_HD void onb(V3 n, V3& v, V3& u) {
	if (fabsf(n.x) > fabsf(n.y))
		v = V3(-n.z, 0.0f, n.x) / sqrtf(n.x * n.x + n.z * n.z);
	else
		v = V3(0.0f, n.z, -n.y) / sqrtf(n.y * n.y + n.z * n.z);
	u = normalize(cross(n, v));
}
// toWorld / toLCS, src/utils/Math.h:612-631
NE_HD V3 to_world(V3 v, V3 ns, V3 ss, V3 ts) {
	return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
}
NE_HD V3 to_lcs(V3 v, V3 ns, V3 ss, V3 ts) { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }

// powerHeuristic, src/utils/Math.h:751-753
NE_HD float power_heuristic(float a, float b) { return (a * a) / (a * a + b * b); }

// convertAreaToSolidAngle, src/utils/Math.h:1112-1126
NE_HD float area_to_solid_angle(float pdfArea, V3 normal, V3 p1, V3 p2) {
	V3 wi = p1 - p2;
	if (length2(wi) == 0) return 0;
	wi = normalize(wi);
	V3 d = p2 - p1;  // glm::distance2(p1,p2) = length2(p2 - p1)
	pdfArea *= dot(d, d) / fabsf(dot(normal, -wi));
	if (isinf(pdfArea)) return 0;
	return pdfArea;
}

// sampleUnitSphere(e1, e2), src/utils/Math.h:442-457
NE_HD V3 sample_unit_sphere(float e1, float e2) {
	float theta = float(NE_TWO_PI * double(e1));
	// acos/sin/cos resolve to the double overloads in the reference build (Math.h is compiled without <math.h>'s
	// float overloads in scope); phi is narrowed to float in between
	float phi = float(acos(double(1.0f - 2.0f * e2)));
	double sp, cp, st, ct;
	sincos(double(phi), &sp, &cp);
	sincos(double(theta), &st, &ct);
	return V3(float(sp * ct), float(sp * st), float(cp));
}

}  // namespace ne
