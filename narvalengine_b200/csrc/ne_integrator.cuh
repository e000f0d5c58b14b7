// ne_integrator.cuh — VolumetricPathIntegrator (integrators/VolumetricPathIntegrator.cpp) restated as STAGES so the
// same code drives both executions:
//   * one thread per path, start to finish (li_path: test hooks, megakernel A/B check) with an ImmediateSink that
//     resolves every next-event query in place, and
//   * the wavefront renderer (ne_wavefront.cu), whose QueueSink turns the two next-event queries of estimateDirect
//     into shadow / transmittance requests that separate kernels resolve and splat.
// Stage map:  classify_hit = Li :187-193 + :244-260 (emission, termination)
//             shade_volume = Li :194-240          shade_surface = Li :262-283
//             estimate_direct = :74-157           sample_one_light = :159-174
//             visibility_tr = :34-72              intersect_tr = :10-32
#pragma once
#include "ne_device.cuh"

namespace ne {

#define NE_MAX_TR_SEGMENTS 64   // guard: the reference's intersectTr can loop forever (DESIGN.md "Deviations")
#define NE_MAX_NULL_SEGMENTS 4096  // guard on Q1 escape/re-enter iterations (the reference has no bound)

// visibilityTr :34-72 — 1 if nothing or an emitter is hit first, else 0 (Q8, Q9).
template <class R, bool FAITHFUL, bool BRICKMAJ>
NE_D float visibility_tr(const DScene& s, V3 p, V3 lightPoint, R& rng, Stats& st) {
	Ray ray;
	ray.o = p;
	ray.d = lightPoint - p;
	Hit h;
	st.shadow_rays++;
	bool hitSurface = intersect_scene(s, ray, h, float(NE_EPSILON3), INFINITY, st);
	if (!hitSurface) return 1.0f;
	int mi = s.inst[h.inst].material;
	if (mi < 0) return 0.0f;
	const DMaterial& m = s.mat[mi];
	if (m.has_light) return 1.0f;
	if (FAITHFUL && m.has_medium && m.volume >= 0)
		(void)grid_tr<R, BRICKMAJ>(s.inst[h.inst], m, s.vol[m.volume], ray, h.tNear, h.tFar, rng, st);  // computed, then discarded by `return 0`
	return 0.0f;
}

// intersectTr :10-32 — marches THROUGH non-medium surfaces until a medium (true, Tr) or nothing (false) (Q12).
// HomogeneousMedia::Tr(ray, hit) = Tr(tFar - tNear) = exp(-extinction * distance * density), materials/HomogeneousMedia.cpp:15-22
NE_D V3 homog_tr(const DMaterial& m, float distance) {
	V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + V3(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
	V3 v = ext * distance * m.density_mult;
	return V3(expf(-v.x), expf(-v.y), expf(-v.z));
}

template <class R, bool FAITHFUL, bool BRICKMAJ>
NE_D bool intersect_tr(const DScene& s, Ray ray, V3& Tr, R& rng, Stats& st) {
	Tr = V3(1.0f);
	if (!FAITHFUL && !s.has_medium) return false;  // can only return true through a medium
	for (int seg = 0; seg < NE_MAX_TR_SEGMENTS; seg++) {
		Hit h;
		st.shadow_rays++;
		bool hitSurface = intersect_scene(s, ray, h, float(NE_EPSILON3), INFINITY, st);
		if (!hitSurface) return false;
		int mi = s.inst[h.inst].material;
		if (mi >= 0 && s.mat[mi].has_medium && s.mat[mi].volume >= 0) {
			if (!FAITHFUL) {  // the wavefront's transmittance requests move the origin to the medium's entry in WCS
				ray.o = ray.at(h.tNear);
				h.tFar -= h.tNear;
				h.tNear = 0;
			}
			Tr = Tr * V3(grid_tr<R, BRICKMAJ>(s.inst[h.inst], s.mat[mi], s.vol[s.mat[mi].volume], ray, h.tNear, h.tFar, rng, st));
			return true;
		}
		if (mi >= 0 && s.mat[mi].has_medium) {  // HomogeneousMedia
			Tr = Tr * homog_tr(s.mat[mi], h.tFar - h.tNear);
			return true;
		}
		ray.o = h.p;
	}
	return false;
}

// Resolves next-event queries in place and sums Ld in the reference's order.
// FAST: medium shading through the hardware-approximation versions (ne_device.cuh "FAST medium shading"): on for production
// arithmetic (the check renderer must take the wavefront's decisions), off for the draw-for-draw tape tests.
template <class R, bool FAITHFUL, bool BRICKMAJ, bool FAST = !FAITHFUL>
struct ImmediateSink {
	static constexpr bool kFast = FAST;
	V3 L;           // radiance of the path so far
	V3 Ld;          // estimateDirect accumulator
	V3 scale;       // unused here: the throughput a queueing sink folds into its requests
	float sel_pdf;  // unused here: light-selection pdf
	NE_D void begin() { Ld = V3(0.0f); }
	NE_D void emit(V3 v) { L = L + v; }
	// light half :100-112: Li *= visibilityTr(p, C); Ld += f * Li * weight / lightPdf
	NE_D void light_term(const DScene& s, V3 p, V3 C, V3 f, V3 Li, float weight, float pdf, R& rng, Stats& st) {
		Li = Li * visibility_tr<R, FAITHFUL, BRICKMAJ>(s, p, C, rng, st);
		if (!is_black(Li)) Ld = Ld + f * Li * weight / pdf;
	}
	// BSDF half :132-153: Ld += f * Li * Tr * weight / scatteringPdf when intersectTr finds a medium
	NE_D void bsdf_term(const DScene& s, Ray ray, V3 f, V3 Li, float weight, float pdf, R& rng, uint32_t stream, Stats& st) {
		V3 Tr;
		Fork<R> fork(rng, stream);
		bool found = intersect_tr<R, FAITHFUL, BRICKMAJ>(s, ray, Tr, fork.get(), st);
		V3 Li2 = found ? Li : V3(0.0f);
		if (!is_black(Li2)) Ld = Ld + f * Li2 * Tr * weight / pdf;
	}
	// uniformSampleOneLight's return value (Ld / lightSelectionPdf), added by the caller as L += T * value
	NE_D V3 end(float selPdf) { return Ld / selPdf; }
};

// estimateDirect :74-157. `lightIdx` = fold index of the chosen light instance; `stream` names the side stream of
// the BSDF-half transmittance walk.
template <int KIND = -1, int LS = -1, class R, class SINK>
NE_D void estimate_direct(const DScene& s, Ray incoming, const Hit& isect, int lightIdx, R& rng, SINK& sink, uint32_t stream, Stats& st) {
	const DInstance& li = s.inst[lightIdx];
	const DMaterial& lm = s.mat[li.material];
	const DInstance& lprim = s.inst[lm.light_owner];
	const DMaterial& m = s.mat[s.inst[isect.inst].material];
	V3 Lrad(lm.li[0], lm.li[1], lm.li[2]);
	constexpr bool FAST = SINK::kFast && KIND == 1;

	Ray wo;
	wo.o = isect.p;
	float lightPdf;
	V3 Li;
	if (LS < 0 && lm.infinite) {
		// InfiniteAreaLight::sampleLi, lights/InfiniteAreaLight.h:58-91. vec2(random(), random()): right to left (A.9)
		const DEnvDist& env = s.env[lm.env];
		float u1 = rng.next(), u0 = rng.next();
		float d0, d1, mapPdf;
		env_sample_continuous(env, u0, u1, d0, d1, mapPdf);
		wo.d = V3(0.0f);
		lightPdf = 0;
		Li = V3(0.0f);
		if (mapPdf != 0) {
			float theta = float(double(d1) * NE_PI), phi = float(double(d0) * 2.0 * NE_PI);
			float cosTheta = cosf(theta), sinTheta = sinf(theta), sinPhi = sinf(phi), cosPhi = cosf(phi);
			V3 polar(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
			V3 up(0.0f, 1.0f, 0.0f), ss, ts;
			onb(up, ss, ts);
			wo.d = to_lcs(polar, up, ss, ts);
			lightPdf = float(double(mapPdf) / (2.0 * NE_PI * NE_PI * double(sinTheta)));
			if (sinTheta == 0) lightPdf = 0;
			V4 t = tex_sample(s.tex[env.tex], d0, d1);
			Li = V3(t.x, t.y, t.z);
		}
	} else if (LS < 0 && lm.directional) {
		// DirectionalLight::sampleLi, lights/DirectionalLight.cpp:13-20: no draws, pdf 1, returns Light::li = 0 (Q23: the
		// light half is dead; wo.d still feeds the medium's phase evaluation of the BSDF half, Q18)
		wo.d = -V3(lm.direction[0], lm.direction[1], lm.direction[2]);
		lightPdf = 1;
		Li = V3(0.0f);
	} else {
		// DiffuseLight::sampleLi, lights/DiffuseLight.cpp:8-20
		V3 A = light_sample_point<LS>(lprim, li, isect, rng);
		wo.d = FAST ? normalize_fast(A - wo.o) : normalize(A - wo.o);  // (a medium uses it for the phase function only)
		lightPdf = light_pdf<LS>(lprim, li, isect, rng);
		Li = Lrad;
	}
	V3 f(0.0f);
	float scatteringPdf = 0;
	const bool isSurface = KIND < 0 ? !m.has_medium : KIND == 0;

	if (lightPdf > 0 && !is_black(Li)) {
		if (isSurface) {
			f = bsdf_eval<KIND>(s, m, incoming.d, wo.d, isect) * fabsf(dot(wo.d, isect.n));
			scatteringPdf = bsdf_pdf<KIND>(s, m, incoming.d, wo.d, isect.n, isect);
		} else {
			f = bsdf_eval<KIND, FAST>(s, m, incoming.d, wo.d, isect);
			scatteringPdf = f.x;
		}
		if (!is_black(f)) {
			V3 C = light_sample_point<LS>(lprim, li, isect, rng);
			sink.light_term(s, isect.p, C, f, Li, power_heuristic(lightPdf, scatteringPdf), lightPdf, rng, st);
		}
	}

	if (isSurface) {
		wo.d = bsdf_sample<KIND>(s, m, incoming.d, isect.n, isect, rng);
		f = bsdf_eval<KIND>(s, m, incoming.d, wo.d, isect);
		f = f * fabsf(dot(wo.d, isect.n));
		scatteringPdf = bsdf_pdf<KIND>(s, m, incoming.d, wo.d, isect.n, isect);
	} else {
		f = bsdf_eval<KIND, FAST>(s, m, incoming.d, wo.d, isect);                            // Q18: f for the light-half direction ...
		wo.d = bsdf_sample<KIND, FAST>(s, m, incoming.d, V3(0.0f, 1.0f, 0.0f), isect, rng);  // ... then a fresh direction
		scatteringPdf = f.x;
	}

	if (!is_black(f) && scatteringPdf > 0) {
		lightPdf = light_pdf<LS>(lprim, li, isect, rng);
		if (lightPdf == 0) return;
		float weight = power_heuristic(scatteringPdf, lightPdf);
		Ray ray;
		ray.o = isect.p;
		ray.d = wo.d;
		sink.bsdf_term(s, ray, f, Lrad, weight, scatteringPdf, rng, stream, st);
	}
}

// uniformSampleOneLight :159-174. Returns what the caller must add as L += T * value (zero for queueing sinks,
// which splat T * value later themselves).
template <int KIND = -1, int LS = -1, class R, class SINK>
NE_D V3 sample_one_light(const DScene& s, Ray incoming, const Hit& isect, R& rng, SINK& sink, uint32_t stream, Stats& st) {
	float r = rng.next();
	if (s.n_lights == 0) return V3(0.0f);  // the reference throws std::out_of_range here
	int i = int(float(s.n_lights) * r);
	float lightPdf = 1.0f / float(s.n_lights);
	(void)rng.next();  // Model::getRandomLightPrimitive (Model.cpp:478-485), result used for the null check
	(void)rng.next();  // second getRandomLightPrimitive call
	sink.begin();
	sink.sel_pdf = lightPdf;
	estimate_direct<KIND, LS>(s, incoming, isect, s.n_models + i, rng, sink, stream, st);
	return sink.end(lightPdf);
}

struct PathState {
	Ray ray;
	V3 T;
	int bounce;
	int guard;      // Q1 escape/re-enter count
	uint32_t nee;   // number of uniformSampleOneLight calls so far (names the side streams)
};
enum { HIT_TERMINATE = 0, HIT_VOLUME = 1, HIT_SURFACE = 2 };
enum { PATH_DONE = 0, PATH_NEXT_BOUNCE = 1, PATH_SAME_BOUNCE = 2 };

// Li :187-193 and :244-260 — what happens right after intersectScene: throughput check, volume / surface split,
// emission for camera rays (Q3), termination on miss / no BSDF.
template <class SINK>
NE_D int classify_hit(const DScene& s, bool did, const Hit& isect, const PathState& ps, SINK& sink) {
	if (is_black(ps.T)) return HIT_TERMINATE;
	int mi = did ? s.inst[isect.inst].material : -1;
	if (did && mi >= 0 && s.mat[mi].has_bsdf && s.mat[mi].transmissive) return HIT_VOLUME;
	if (ps.bounce == 0 && did && mi >= 0 && s.mat[mi].has_light) {
		const DMaterial& m = s.mat[mi];
		sink.emit(ps.T * V3(m.li[0], m.li[1], m.li[2]));
	}
	// else at bounce 0 (:249-256): the sum of Light::Le over every light model - 0 for DiffuseLight (lights/Light.h:20-22),
	// le for DirectionalLight (lights/DirectionalLight.cpp:4-6)
	else if (ps.bounce == 0 && (s.n_directional | s.n_infinite)) {
		for (int i = s.n_models; i < s.n_inst; i++) {
			const DMaterial& lm = s.mat[s.inst[i].material];
			if (lm.directional) sink.emit(ps.T * V3(lm.li[0], lm.li[1], lm.li[2]));
			if (lm.infinite) sink.emit(ps.T * env_le(s, s.env[lm.env], s.inst[i].Mi, ps.ray.d));  // InfiniteAreaLight::Le :47-56
		}
	}
	if (!did || mi < 0 || !s.mat[mi].has_bsdf) return HIT_TERMINATE;
	return HIT_SURFACE;
}

// Li :194-240 — the volume branch, in three pieces so the wavefront can run the tracking loop in its own kernel.
// (1) :198-201 move the origin to the volume's boundary (idempotent once tNear is 0).
NE_D void volume_enter(PathState& ps, Hit& isect) {
	ps.ray.o = ps.ray.at(isect.tNear);
	isect.tFar = isect.tFar - isect.tNear;
	isect.tNear = 0;
}
// (2a) :209-213 escape (Q1: detected by value in the reference; does not consume a bounce).
NE_D int volume_escape(PathState& ps, const Hit& isect) {
	ps.ray.o = ps.ray.at(isect.tFar + 0.01f);
	if (++ps.guard > NE_MAX_NULL_SEGMENTS) return PATH_DONE;
	return PATH_SAME_BOUNCE;
}
template <int LS = -1, class R, class SINK>
NE_D int volume_collision(const DScene& s, PathState& ps, const Hit& isect, Ray scattered, V3 a, R& rng, SINK& sink, Stats& st);
// (2b) :215-236 real collision at parameter t of the OCS ray `rayO`.
template <int LS = -1, class R, class SINK>
NE_D int volume_scatter(const DScene& s, PathState& ps, const Hit& isect, const Ray& rayO, float t, R& rng, SINK& sink, Stats& st) {
	const DInstance& in = s.inst[isect.inst];
	const DMaterial& m = s.mat[in.material];
	Ray scattered = grid_scatter<SINK::kFast>(s, in, m, rayO, t, isect, rng, st);
	V3 sc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
	V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + sc;
	V3 a = sc / ext;
	return volume_collision<LS>(s, ps, isect, scattered, a, rng, sink, st);
}
// Li :209-236 once the medium has answered with `a` and `scattered`.
template <int LS, class R, class SINK>
NE_D int volume_collision(const DScene& s, PathState& ps, const Hit& isect, Ray scattered, V3 a, R& rng, SINK& sink, Stats& st) {
	const DMaterial& m = s.mat[s.inst[isect.inst].material];
	if (all_one(a)) return volume_escape(ps, isect);  // Q1 / Q1b: an escape is recognised by the value (1,1,1)
	ps.T = ps.T * a;
	V3 Tnew;
	if (SINK::kFast) {
		// a phase function is its own pdf: fr / pdf is x / x = 1 in IEEE arithmetic too, unless x is 0, inf or NaN
		float ph = phase_eval_fast(m, ps.ray.d, scattered.d);
		if (!(ph > 0.0f) || isinf(ph)) return PATH_DONE;
		Tnew = ps.T;
	} else {
		V3 phaseFr = bsdf_eval<1>(s, m, ps.ray.d, scattered.d, isect);
		float phasePdf = bsdf_pdf<1>(s, m, ps.ray.d, scattered.d, isect.n, isect);
		if (is_black(phaseFr) || phasePdf == 0.f) return PATH_DONE;
		Tnew = ps.T * (phaseFr / phasePdf);
	}
	sink.scale = Tnew;  // L += T * lightSample happens AFTER the throughput update (:228-232)
	V3 lightSample = sample_one_light<1, LS>(s, scattered, isect, rng, sink, 1u + ps.nee++, st);  // Q2
	ps.T = Tnew;
	sink.emit(ps.T * lightSample);
	ps.ray = scattered;
	return PATH_NEXT_BOUNCE;
}
// The volume branch for a HomogeneousMedia (materials/HomogeneousMedia.cpp:24-51): closed-form free flight in WCS
// (no OCS transform), `t` divided by the segment length (sic), phase sample about the proxy box's hit normal; the
// escape value Tr / avg(Tr) is (1,1,1) only when the rounding of (x+x+x)/3 allows it - otherwise Li treats the
// escape as a collision that keeps its direction.
template <class R, class SINK>
NE_D int shade_volume_homog(const DScene& s, PathState& ps, Hit& isect, R& rng, SINK& sink, Stats& st) {
	const DMaterial& m = s.mat[s.inst[isect.inst].material];
	volume_enter(ps, isect);
	V3 sc(m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]);
	V3 ext = V3(m.sigma_a[0], m.sigma_a[1], m.sigma_a[2]) + sc;
	float t = -logf(1 - rng.next()) / avg(ext);
	float dist = isect.tFar - isect.tNear;
	t = t / dist;
	bool sampled = t < dist;
	Ray scattered;
	if (sampled) {
		st.scatter_events++;
		scattered.o = ps.ray.at(t);
		scattered.d = bsdf_sample<1, SINK::kFast>(s, m, ps.ray.d, isect.n, isect, rng);
	} else {
		scattered.o = ps.ray.at(dist + 0.001f);
		scattered.d = ps.ray.d;
	}
	V3 Tr = homog_tr(m, t);
	V3 density = sampled ? (ext * Tr) : Tr;
	float pdf = avg(density);
	if (pdf == 0) pdf = 1;
	V3 a = sampled ? (Tr * sc / pdf) : Tr / pdf;
	return volume_collision(s, ps, isect, scattered, a, rng, sink, st);
}
// The whole volume branch in one go (one thread per path).
template <class R, bool BRICKMAJ, class SINK>
NE_D int shade_volume(const DScene& s, PathState& ps, Hit& isect, R& rng, SINK& sink, Stats& st) {
	const DInstance& in = s.inst[isect.inst];
	const DMaterial& m = s.mat[in.material];
	if (m.volume < 0) return shade_volume_homog(s, ps, isect, rng, sink, st);
	const DVolume& v = s.vol[m.volume];
	volume_enter(ps, isect);
	Ray rayO = transform_ray(ps.ray, in.Mi);
	typename WalkRngOf<BRICKMAJ, R>::type wr;
	wr.start(rng);
	Tracker<BRICKMAJ> trk;
	trk.init(v, m, rayO, 0.0f, isect.tFar, wr, st);
	if (delta_walk(v, trk, wr, st, NE_NO_BUDGET) != TRACK_CANDIDATE) return volume_escape(ps, isect);
	return volume_scatter(s, ps, isect, rayO, trk.t, rng, sink, st);
}

// Li :262-283 — the surface branch.
template <class R, int LS = -1, class SINK>
NE_D int shade_surface(const DScene& s, PathState& ps, const Hit& isect, R& rng, SINK& sink, Stats& st) {
	const DMaterial& m = s.mat[s.inst[isect.inst].material];
	st.surface_events++;
	sink.scale = ps.T;
	V3 ls = sample_one_light<0, LS>(s, ps.ray, isect, rng, sink, 1u + ps.nee++, st);
	sink.emit(ps.T * ls);
	Ray scattered;
	scattered.o = isect.p;
	scattered.d = bsdf_sample<0>(s, m, ps.ray.d, isect.n, isect, rng);
	float bsdfPdf = bsdf_pdf<0>(s, m, ps.ray.d, scattered.d, isect.n, isect);
	V3 fr = bsdf_eval<0>(s, m, ps.ray.d, scattered.d, isect);
	if (is_black(fr) || bsdfPdf == 0.f) return PATH_DONE;
	ps.T = ps.T * (fr * fabsf(dot(ps.ray.d, isect.n)) / bsdfPdf);  // Q5
	ps.ray = scattered;
	return PATH_NEXT_BOUNCE;
}

// Li :176-301, one thread start to finish.
template <class R, bool FAITHFUL, bool BRICKMAJ, bool FAST = !FAITHFUL>
NE_D V3 li_path(const DScene& s, Ray incoming, int bounces, R& rng, Stats& st) {
	ImmediateSink<R, FAITHFUL, BRICKMAJ, FAST> sink;
	sink.L = V3(0.0f);
	PathState ps;
	ps.ray = incoming;
	ps.T = V3(1.0f);
	ps.bounce = 0;
	ps.guard = 0;
	ps.nee = 0;
	Hit isect;
	while (ps.bounce < bounces) {
		st.extend_rays++;
		bool did = intersect_scene(s, ps.ray, isect, float(NE_EPSILON12), INFINITY, st);
		int kind = classify_hit(s, did, isect, ps, sink);
		if (kind == HIT_TERMINATE) break;
		int next = kind == HIT_VOLUME ? shade_volume<R, BRICKMAJ>(s, ps, isect, rng, sink, st) : shade_surface<R>(s, ps, isect, rng, sink, st);
		if (next == PATH_DONE) break;
		if (next == PATH_NEXT_BOUNCE) ps.bounce++;
	}
	return sink.L;
}

}  // namespace ne
