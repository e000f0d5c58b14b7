// ne_frontend.cpp — the scene front end: what io/SceneReader.cpp:10-675 and
// core/ResourceManager.cpp:165-315 do for the reference, producing the POD ne_b200_scene_desc the backend uploads.
// Pure host code with no third-party dependency except zlib (PNG): its own JSON reader (the reference uses rapidjson),
// `.vol` reader/writer (ResourceManager::loadVolasTexture :222-286, incl. its space-terminated-token parser), OBJ reader
// and glTF 2.0 / .glb reader (the reference goes through assimp with Triangulate | FlipUVs and NO vertex joining, ResourceManager.cpp:59: one
// vertex per face corner, so a triangle's three vertices are consecutive - which is what Triangle::samplePointOnTexture
// assumes, Q30), PNG reader (stbi_load(..., STBI_rgb_alpha), ResourceManager.cpp:288-315) and the framebuffer
// consumers (§8f rank 2): PNG / EXR as materials/Texture.h:44-75 saveImage writes them, and OfflineEngine::coreLoop's
// 16-bit output.ppm (core/OfflineEngine.cpp:78-139).
// Everything the reference LOG(FATAL)s on returns NE_B200_ERR_INVALID with a message instead.
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "ne_b200.h"

namespace ne {
void set_error(const std::string& s);  // ne_api.cu
}
using ne::set_error;

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Minimal JSON (RFC 8259) document: objects keep insertion order; unknown keys are simply never looked up.
// ---------------------------------------------------------------------------------------------------------------
struct JValue {
	enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
	bool b = false;
	double num = 0;
	std::string str;
	std::vector<JValue> arr;
	std::vector<std::pair<std::string, JValue>> obj;
	const JValue* get(const char* key) const {
		if (kind != Object) return nullptr;
		for (const auto& kv : obj)
			if (kv.first == key) return &kv.second;
		return nullptr;
	}
	bool has(const char* key) const { return get(key) != nullptr; }
};

struct JParser {
	const char* p;
	const char* end;
	std::string err;
	void ws() {
		while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
	}
	bool fail(const std::string& m) {
		if (err.empty()) err = m;
		return false;
	}
	bool parseString(std::string& out) {
		if (p >= end || *p != '"') return fail("expected string");
		p++;
		out.clear();
		while (p < end && *p != '"') {
			if (*p == '\\') {
				if (++p >= end) return fail("bad escape");
				switch (*p) {
				case '"': out += '"'; break;
				case '\\': out += '\\'; break;
				case '/': out += '/'; break;
				case 'b': out += '\b'; break;
				case 'f': out += '\f'; break;
				case 'n': out += '\n'; break;
				case 'r': out += '\r'; break;
				case 't': out += '\t'; break;
				case 'u': {
					if (end - p < 5) return fail("bad \\u escape");
					unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
					p += 4;
					if (cp < 0x80) out += char(cp);
					else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
					else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
					break;
				}
				default: return fail("bad escape");
				}
				p++;
			} else {
				out += *p++;
			}
		}
		if (p >= end) return fail("unterminated string");
		p++;
		return true;
	}
	bool parse(JValue& v, int depth = 0) {
		if (depth > 64) return fail("nesting too deep");
		ws();
		if (p >= end) return fail("unexpected end of input");
		if (*p == '{') {
			v.kind = JValue::Object;
			p++;
			ws();
			if (p < end && *p == '}') { p++; return true; }
			while (true) {
				ws();
				std::string key;
				if (!parseString(key)) return false;
				ws();
				if (p >= end || *p != ':') return fail("expected ':'");
				p++;
				JValue child;
				if (!parse(child, depth + 1)) return false;
				v.obj.emplace_back(std::move(key), std::move(child));
				ws();
				if (p < end && *p == ',') { p++; continue; }
				if (p < end && *p == '}') { p++; return true; }
				return fail("expected ',' or '}'");
			}
		}
		if (*p == '[') {
			v.kind = JValue::Array;
			p++;
			ws();
			if (p < end && *p == ']') { p++; return true; }
			while (true) {
				JValue child;
				if (!parse(child, depth + 1)) return false;
				v.arr.push_back(std::move(child));
				ws();
				if (p < end && *p == ',') { p++; continue; }
				if (p < end && *p == ']') { p++; return true; }
				return fail("expected ',' or ']'");
			}
		}
		if (*p == '"') {
			v.kind = JValue::String;
			return parseString(v.str);
		}
		if (!strncmp(p, "true", std::min<size_t>(4, end - p)) && end - p >= 4) { v.kind = JValue::Bool; v.b = true; p += 4; return true; }
		if (!strncmp(p, "false", std::min<size_t>(5, end - p)) && end - p >= 5) { v.kind = JValue::Bool; v.b = false; p += 5; return true; }
		if (!strncmp(p, "null", std::min<size_t>(4, end - p)) && end - p >= 4) { v.kind = JValue::Null; p += 4; return true; }
		char* e = nullptr;
		double d = strtod(p, &e);
		if (e == p) return fail("unexpected character");
		v.kind = JValue::Number;
		v.num = d;
		p = e;
		return true;
	}
};

// ---------------------------------------------------------------------------------------------------------------
// Assets
// ---------------------------------------------------------------------------------------------------------------
bool read_file(const std::string& path, std::string& out) {
	std::ifstream f(path, std::ios::binary);
	if (!f) return false;
	std::ostringstream ss;
	ss << f.rdbuf();
	out = ss.str();
	return true;
}

// ResourceManager::loadVolasTexture, core/ResourceManager.cpp:222-286. Line 1: numbers EACH followed by one space
// (only space-terminated tokens are captured); line 2 is read and discarded; the rest: the lines are concatenated with
// '\n' and split on single spaces, each token through std::stof (which stops at the first character that is not
// part of a number, so "0.5\n0.6" yields 0.5 and drops 0.6).
int vol_parse(const std::string& text, int32_t dims[3], std::vector<float>* grid) {
	size_t l1 = text.find('\n');
	std::string first = text.substr(0, l1);
	if (!first.empty() && first.back() == '\r') first.pop_back();
	float res[3] = {0, 0, 0};
	int count = 0;
	size_t start = 0, endp = first.find(' ');
	while (endp != std::string::npos) {
		if (count >= 3) break;
		std::string tok = first.substr(start, endp - start);
		char* e = nullptr;
		float v = strtof(tok.c_str(), &e);
		if (e == tok.c_str()) { set_error(".vol: bad resolution token '" + tok + "'"); return NE_B200_ERR_INVALID; }
		res[count++] = v;
		start = endp + 1;
		endp = first.find(' ', start);
	}
	if (count < 3 || !(res[0] >= 1) || !(res[1] >= 1) || !(res[2] >= 1)) { set_error(".vol: the first line must hold 'W H D ' (each number followed by a space)"); return NE_B200_ERR_INVALID; }
	// untrusted input: each side fits an int32 comfortably, and the grid cannot hold more voxels than the file has
	// tokens (every value is at least one character and one space)
	if (!(res[0] <= 65536.0f) || !(res[1] <= 65536.0f) || !(res[2] <= 65536.0f)) { set_error(".vol: resolution out of range (more than 65536 on a side)"); return NE_B200_ERR_INVALID; }
	dims[0] = int32_t(res[0]); dims[1] = int32_t(res[1]); dims[2] = int32_t(res[2]);
	if (!grid) return NE_B200_OK;
	size_t n = size_t(dims[0]) * dims[1] * dims[2];
	if (n > (size_t(1) << 36)) { set_error(".vol: W*H*D exceeds 2^36 voxels"); return NE_B200_ERR_INVALID; }
	grid->assign(n, 0.0f);
	if (l1 == std::string::npos) return NE_B200_OK;
	size_t l2 = text.find('\n', l1 + 1);  // second line: discarded
	if (l2 == std::string::npos) return NE_B200_OK;
	const char* p = text.c_str() + l2 + 1;
	const char* endt = text.c_str() + text.size();
	size_t k = 0;
	while (p < endt) {
		const char* sp = static_cast<const char*>(memchr(p, ' ', endt - p));
		if (!sp) break;  // a last token without a trailing space is not captured
		char* e = nullptr;
		float v = strtof(p, &e);
		if (e == p || e > sp) {
			if (e == p) { set_error(".vol: empty or non-numeric density token"); return NE_B200_ERR_INVALID; }
		}
		if (k >= n) { set_error(".vol: more density values than W*H*D"); return NE_B200_ERR_INVALID; }
		(*grid)[k++] = v;
		p = sp + 1;
	}
	return NE_B200_OK;
}

// PNG -> RGBA8 like stbi_load(path, &w, &h, &c, STBI_rgb_alpha): 8/16-bit grey, grey+alpha, RGB, RGBA, palette
// (with tRNS), non-interlaced. Row 0 = top row (no flip), as ResourceManager::loadTexture leaves it.
bool png_decode(const std::string& data, int& w, int& h, std::vector<uint8_t>& rgba, std::string& err) {
	static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	if (data.size() < 8 || memcmp(data.data(), sig, 8)) { err = "not a PNG file"; return false; }
	auto be32 = [&](size_t o) { return (uint32_t(uint8_t(data[o])) << 24) | (uint32_t(uint8_t(data[o + 1])) << 16) | (uint32_t(uint8_t(data[o + 2])) << 8) | uint32_t(uint8_t(data[o + 3])); };
	size_t o = 8;
	int depth = 0, ctype = 0, interlace = 0;
	std::string idat;
	std::vector<uint8_t> plte, trns;
	bool haveHdr = false;
	while (o + 8 <= data.size()) {
		uint32_t len = be32(o);
		std::string type = data.substr(o + 4, 4);
		if (o + 12 + size_t(len) > data.size()) { err = "truncated PNG chunk"; return false; }
		const char* body = data.data() + o + 8;
		if (type == "IHDR") {
			if (len < 13) { err = "bad IHDR"; return false; }
			w = int(be32(o + 8)); h = int(be32(o + 12));
			depth = uint8_t(body[8]); ctype = uint8_t(body[9]); interlace = uint8_t(body[12]);
			haveHdr = true;
		} else if (type == "PLTE") plte.assign(body, body + len);
		else if (type == "tRNS") trns.assign(body, body + len);
		else if (type == "IDAT") idat.append(body, len);
		else if (type == "IEND") break;
		o += 12 + size_t(len);
	}
	if (!haveHdr || w <= 0 || h <= 0) { err = "PNG without IHDR"; return false; }
	if (w > 65536 || h > 65536 || size_t(w) * size_t(h) > (size_t(1) << 28)) { err = "PNG larger than 65536 on a side or 2^28 pixels"; return false; }
	if (interlace) { err = "interlaced PNG is not supported"; return false; }
	int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!ch || (depth != 8 && depth != 16 && !(ctype == 3 || ctype == 0))) { err = "unsupported PNG colour type / depth"; return false; }
	if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) { err = "unsupported PNG bit depth"; return false; }
	size_t bpp = std::max<size_t>(1, size_t(ch) * depth / 8);        // bytes per complete pixel (filter unit)
	size_t stride = (size_t(w) * ch * depth + 7) / 8;
	std::vector<uint8_t> raw((stride + 1) * size_t(h));
	uLongf rawLen = uLongf(raw.size());
	int zr = uncompress(raw.data(), &rawLen, reinterpret_cast<const Bytef*>(idat.data()), uLong(idat.size()));
	if (zr != Z_OK || rawLen != raw.size()) { err = "PNG inflate failed"; return false; }
	std::vector<uint8_t> prev(stride, 0), cur(stride);
	rgba.assign(size_t(w) * h * 4, 255);
	for (int y = 0; y < h; y++) {
		const uint8_t* row = &raw[(stride + 1) * size_t(y)];
		int ft = row[0];
		for (size_t i = 0; i < stride; i++) {
			int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = row[1 + i], v;
			switch (ft) {
			case 0: v = x; break;
			case 1: v = x + a; break;
			case 2: v = x + b; break;
			case 3: v = x + ((a + b) >> 1); break;
			case 4: { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); v = x + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)); break; }
			default: err = "bad PNG filter"; return false;
			}
			cur[i] = uint8_t(v);
		}
		for (int x = 0; x < w; x++) {
			uint8_t* px = &rgba[(size_t(y) * w + x) * 4];
			auto sample = [&](int k) -> int {  // k-th channel sample of pixel x, scaled to 8 bits
				if (depth == 8) return cur[size_t(x) * ch + k];
				if (depth == 16) return cur[(size_t(x) * ch + k) * 2];  // stb keeps the high byte
				int bit = (x * ch + k) * depth;
				int v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
				return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
			};
			if (ctype == 0) { px[0] = px[1] = px[2] = uint8_t(sample(0)); }
			else if (ctype == 4) { px[0] = px[1] = px[2] = uint8_t(sample(0)); px[3] = uint8_t(sample(1)); }
			else if (ctype == 2) { px[0] = uint8_t(sample(0)); px[1] = uint8_t(sample(1)); px[2] = uint8_t(sample(2)); }
			else if (ctype == 6) { px[0] = uint8_t(sample(0)); px[1] = uint8_t(sample(1)); px[2] = uint8_t(sample(2)); px[3] = uint8_t(sample(3)); }
			else {
				size_t idx = size_t(sample(0));
				if (idx * 3 + 2 < plte.size()) { px[0] = plte[idx * 3]; px[1] = plte[idx * 3 + 1]; px[2] = plte[idx * 3 + 2]; }
				if (idx < trns.size()) px[3] = trns[idx];
			}
		}
		prev.swap(cur);
	}
	return true;
}

uint32_t crc_of(const std::string& type, const std::vector<uint8_t>& body) {
	uLong c = crc32(0L, reinterpret_cast<const Bytef*>(type.data()), 4);
	if (!body.empty()) c = crc32(c, body.data(), uInt(body.size()));
	return uint32_t(c);
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void png_chunk(std::vector<uint8_t>& out, const std::string& type, const std::vector<uint8_t>& body) {
	put_be32(out, uint32_t(body.size()));
	out.insert(out.end(), type.begin(), type.end());
	out.insert(out.end(), body.begin(), body.end());
	put_be32(out, crc_of(type, body));
}

// OBJ: v / vt / f (with negative indices, v/vt/vn forms); polygons are fan-triangulated (aiProcess_Triangulate on
// convex faces); one vertex per face corner, V flipped (aiProcess_FlipUVs).
int obj_parse(const std::string& text, std::vector<float>& pos, std::vector<float>& uv, std::vector<uint32_t>& idx, bool& hasUv) {
	std::vector<float> v, vt;
	hasUv = false;
	std::istringstream in(text);
	std::string line;
	std::vector<std::pair<int, int>> corners;
	while (std::getline(in, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		const char* s = line.c_str();
		while (*s == ' ' || *s == '\t') s++;
		if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
			float x = 0, y = 0, z = 0;
			if (sscanf(s + 1, "%f %f %f", &x, &y, &z) < 3) { set_error("obj: bad 'v' line"); return NE_B200_ERR_INVALID; }
			v.push_back(x); v.push_back(y); v.push_back(z);
		} else if (s[0] == 'v' && s[1] == 't' && (s[2] == ' ' || s[2] == '\t')) {
			float a = 0, b = 0;
			if (sscanf(s + 2, "%f %f", &a, &b) < 1) { set_error("obj: bad 'vt' line"); return NE_B200_ERR_INVALID; }
			vt.push_back(a); vt.push_back(b);
		} else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
			corners.clear();
			const char* q = s + 1;
			while (*q) {
				while (*q == ' ' || *q == '\t') q++;
				if (!*q) break;
				char* e = nullptr;
				long vi = strtol(q, &e, 10), ti = 0;
				if (e == q) { set_error("obj: bad face index"); return NE_B200_ERR_INVALID; }
				q = e;
				if (*q == '/') {
					q++;
					if (*q != '/') { ti = strtol(q, &e, 10); q = e; }
					if (*q == '/') { q++; strtol(q, &e, 10); q = e; }
				}
				long nv = long(v.size() / 3), nt = long(vt.size() / 2);
				vi = vi < 0 ? nv + vi : vi - 1;
				ti = ti < 0 ? nt + ti : ti - 1;
				if (vi < 0 || vi >= nv) { set_error("obj: vertex index out of range"); return NE_B200_ERR_INVALID; }
				if (ti >= nt) { set_error("obj: texture index out of range"); return NE_B200_ERR_INVALID; }
				corners.emplace_back(int(vi), int(ti));
			}
			for (size_t k = 2; k < corners.size(); k++) {
				const std::pair<int, int> tri[3] = {corners[0], corners[k - 1], corners[k]};
				for (const auto& c : tri) {
					idx.push_back(uint32_t(pos.size() / 3));
					pos.push_back(v[3 * c.first]); pos.push_back(v[3 * c.first + 1]); pos.push_back(v[3 * c.first + 2]);
					if (c.second >= 0) { uv.push_back(vt[2 * c.second]); uv.push_back(1.0f - vt[2 * c.second + 1]); hasUv = true; }
					else { uv.push_back(0.0f); uv.push_back(0.0f); }
				}
			}
		}
	}
	return NE_B200_OK;
}

// glTF 2.0 (.gltf with external / base64 buffers, or .glb): what the reference gets from assimp for a `gltf` primitive and
// Model::processNode keeps of it (primitives/Model.cpp:323-335): the meshes of the node tree in depth-first order, one
// after the other, WITHOUT the nodes' transforms; triangles only; indexed vertices as stored; V flipped (aiProcess_FlipUVs).
bool base64_decode(const std::string& in, std::string& out) {
	static int8_t T[256];
	static bool init = false;
	if (!init) {
		for (int i = 0; i < 256; i++) T[i] = -1;
		const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
		for (int i = 0; i < 64; i++) T[uint8_t(a[i])] = int8_t(i);
		init = true;
	}
	uint32_t acc = 0;
	int bits = 0;
	for (char c : in) {
		if (c == '=' || c == '\n' || c == '\r') continue;
		int v = T[uint8_t(c)];
		if (v < 0) return false;
		acc = (acc << 6) | uint32_t(v);
		bits += 6;
		if (bits >= 8) { bits -= 8; out += char((acc >> bits) & 0xFF); }
	}
	return true;
}

int gltf_parse(const std::string& file, const std::string& dir, std::vector<float>& pos, std::vector<float>& uv, std::vector<uint32_t>& idx, bool& hasUv) {
	auto bad = [&](const std::string& e) { set_error("gltf: " + e); return NE_B200_ERR_INVALID; };
	std::string json = file, glbBin;
	if (file.size() >= 20 && !memcmp(file.data(), "glTF", 4)) {  // .glb: 12-byte header, JSON chunk, optional BIN chunk
		auto le32 = [&](size_t o) { uint32_t v; memcpy(&v, file.data() + o, 4); return v; };
		size_t o = 12;
		json.clear();
		while (o + 8 <= file.size()) {
			uint32_t len = le32(o), type = le32(o + 4);
			if (o + 8 + size_t(len) > file.size()) return bad("truncated glb chunk");
			if (type == 0x4E4F534A) json.assign(file.data() + o + 8, len);
			else if (type == 0x004E4942) glbBin.assign(file.data() + o + 8, len);
			o += 8 + size_t(len);
		}
	}
	JParser jp{json.c_str(), json.c_str() + json.size(), ""};
	JValue doc;
	if (!jp.parse(doc) || doc.kind != JValue::Object) return bad("malformatted json: " + jp.err);
	auto arr = [&](const char* k) -> const std::vector<JValue>* { const JValue* v = doc.get(k); return (v && v->kind == JValue::Array) ? &v->arr : nullptr; };
	const auto* buffers = arr("buffers");
	const auto* views = arr("bufferViews");
	const auto* accessors = arr("accessors");
	const auto* meshes = arr("meshes");
	const auto* nodes = arr("nodes");
	if (!buffers || !views || !accessors || !meshes) return bad("missing buffers / bufferViews / accessors / meshes");
	std::vector<std::string> data(buffers->size());
	for (size_t i = 0; i < buffers->size(); i++) {
		const JValue* uri = (*buffers)[i].get("uri");
		if (!uri) { data[i] = glbBin; continue; }
		if (uri->str.compare(0, 5, "data:") == 0) {
			size_t c = uri->str.find("base64,");
			if (c == std::string::npos || !base64_decode(uri->str.substr(c + 7), data[i])) return bad("bad data: URI");
		} else if (!read_file(dir + uri->str, data[i])) return bad("couldn't read the file at " + dir + uri->str);
	}
	auto num = [](const JValue& o, const char* k, double dflt) { const JValue* v = o.get(k); return (v && v->kind == JValue::Number) ? v->num : dflt; };
	struct View { const uint8_t* p; size_t stride, count; int comp, n; bool normalized; };
	auto view_of = [&](int ai, View& out) -> bool {
		if (ai < 0 || size_t(ai) >= accessors->size()) return false;
		const JValue& a = (*accessors)[ai];
		int bv = int(num(a, "bufferView", -1));
		if (bv < 0 || size_t(bv) >= views->size()) return false;
		const JValue& v = (*views)[bv];
		int buf = int(num(v, "buffer", -1));
		if (buf < 0 || size_t(buf) >= data.size()) return false;
		const JValue* type = a.get("type");
		if (!type) return false;
		out.n = type->str == "SCALAR" ? 1 : type->str == "VEC2" ? 2 : type->str == "VEC3" ? 3 : type->str == "VEC4" ? 4 : 0;
		out.comp = int(num(a, "componentType", 0));
		size_t cs = out.comp == 5126 || out.comp == 5125 ? 4 : (out.comp == 5123 || out.comp == 5122) ? 2 : (out.comp == 5121 || out.comp == 5120) ? 1 : 0;
		if (!out.n || !cs) return false;
		// untrusted numbers: negative or huge values are rejected before they become sizes, and the bound check below is
		// done in a form that cannot wrap around
		const double dCount = num(a, "count", 0), dStride = num(v, "byteStride", 0), dOff = num(v, "byteOffset", 0) + num(a, "byteOffset", 0);
		const double lim = double(data[buf].size());
		if (!(dCount >= 0 && dCount <= lim) || !(dStride >= 0 && dStride <= 65536) || !(dOff >= 0 && dOff <= lim)) return false;
		out.count = size_t(dCount);
		out.stride = size_t(dStride);
		if (!out.stride) out.stride = cs * out.n;
		const JValue* nz = a.get("normalized");
		out.normalized = nz && nz->kind == JValue::Bool && nz->b;
		size_t off = size_t(dOff);
		const size_t elem = cs * out.n, size = data[buf].size();
		if (out.count && (elem > size || off > size - elem || (out.count - 1) > (size - elem - off) / out.stride)) return false;
		out.p = reinterpret_cast<const uint8_t*>(data[buf].data()) + off;
		return true;
	};
	auto fetch = [](const View& v, size_t i, int c) -> double {
		const uint8_t* q = v.p + i * v.stride;
		switch (v.comp) {
		case 5126: { float f; memcpy(&f, q + 4 * c, 4); return f; }
		case 5125: { uint32_t u; memcpy(&u, q + 4 * c, 4); return u; }
		case 5123: { uint16_t u; memcpy(&u, q + 2 * c, 2); return v.normalized ? u / 65535.0 : u; }
		case 5121: return v.normalized ? q[c] / 255.0 : q[c];
		default: return 0;
		}
	};
	hasUv = false;
	auto add_mesh = [&](int mi) -> int {
		if (mi < 0 || size_t(mi) >= meshes->size()) return bad("mesh index out of range");
		const JValue* prims = (*meshes)[mi].get("primitives");
		if (!prims || prims->kind != JValue::Array) return NE_B200_OK;
		for (const JValue& pr : prims->arr) {
			if (int(num(pr, "mode", 4)) != 4) continue;  // triangles only
			const JValue* at = pr.get("attributes");
			if (!at) continue;
			View vp, vt, vi;
			if (!view_of(int(num(*at, "POSITION", -1)), vp) || vp.n != 3 || vp.comp != 5126) return bad("bad POSITION accessor");
			bool uvOk = at->has("TEXCOORD_0") && view_of(int(num(*at, "TEXCOORD_0", -1)), vt) && vt.n == 2;
			const uint32_t base = uint32_t(pos.size() / 3);
			for (size_t i = 0; i < vp.count; i++) {
				for (int c = 0; c < 3; c++) pos.push_back(float(fetch(vp, i, c)));
				if (uvOk && i < vt.count) { uv.push_back(float(fetch(vt, i, 0))); uv.push_back(1.0f - float(fetch(vt, i, 1))); hasUv = true; }
				else { uv.push_back(0.0f); uv.push_back(0.0f); }
			}
			if (pr.has("indices")) {
				if (!view_of(int(num(pr, "indices", -1)), vi) || vi.n != 1) return bad("bad indices accessor");
				for (size_t i = 0; i + 2 < vi.count; i += 3)
					for (int k = 0; k < 3; k++) {
						uint32_t ix = uint32_t(fetch(vi, i + k, 0));
						if (ix >= vp.count) return bad("index out of range");
						idx.push_back(base + ix);
					}
			} else {
				for (size_t i = 0; i + 2 < vp.count; i += 3)
					for (int k = 0; k < 3; k++) idx.push_back(base + uint32_t(i + k));
			}
		}
		return NE_B200_OK;
	};
	// depth-first over the scene's node tree, like Model::processNode
	std::vector<int> stack;
	const auto* scenesArr = arr("scenes");
	if (nodes && scenesArr && !scenesArr->empty()) {
		size_t si = size_t(num(doc, "scene", 0));
		if (si >= scenesArr->size()) si = 0;
		const JValue* roots = (*scenesArr)[si].get("nodes");
		if (roots && roots->kind == JValue::Array)
			for (size_t i = roots->arr.size(); i-- > 0;) stack.push_back(int(roots->arr[i].num));
	} else if (!nodes) {
		for (size_t m = 0; m < meshes->size(); m++) { int rc = add_mesh(int(m)); if (rc) return rc; }
	}
	size_t visited = 0;
	while (!stack.empty()) {
		int ni = stack.back();
		stack.pop_back();
		if (ni < 0 || size_t(ni) >= nodes->size() || ++visited > 100000) return bad("bad node graph");
		const JValue& n = (*nodes)[ni];
		if (n.has("mesh")) { int rc = add_mesh(int(num(n, "mesh", -1))); if (rc) return rc; }
		const JValue* ch = n.get("children");
		if (ch && ch->kind == JValue::Array)
			for (size_t i = ch->arr.size(); i-- > 0;) stack.push_back(int(ch->arr[i].num));
	}
	return NE_B200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// The loaded scene: owns every array the descriptor points to.
// ---------------------------------------------------------------------------------------------------------------
struct ne_b200_scene_file {
	ne_b200_scene_desc desc{};
	std::vector<ne_b200_texture> textures;
	std::vector<ne_b200_volume> volumes;
	std::vector<ne_b200_material> materials;
	std::vector<ne_b200_primitive> primitives;
	std::vector<std::unique_ptr<std::vector<uint8_t>>> bytes;
	std::vector<std::unique_ptr<std::vector<float>>> floats;
	std::vector<std::unique_ptr<std::vector<uint32_t>>> uints;
	std::map<std::string, int> materialByName;
	std::map<std::string, int> volumeByPath, imageByPath;
	float camPosition[3] = {0, 0, 0}, camLookAt[3] = {0, 0, 0};
	float vfov = 45, focus = 1;
	ne_b200_render_settings settings{};
	std::string resources;

	const float* keepF(std::vector<float>&& v) {
		floats.emplace_back(new std::vector<float>(std::move(v)));
		return floats.back()->data();
	}
	int addConstTexture(const float* v, int n) {  // Texture(1, 1, R32F / RGB32F, clamp)
		ne_b200_texture t{};
		t.width = t.height = 1;
		t.format = n == 1 ? NE_B200_TEX_R32F : NE_B200_TEX_RGB32F;
		t.wrap_u = t.wrap_v = NE_B200_WRAP_CLAMP;
		t.texels = keepF(std::vector<float>(v, v + n));
		textures.push_back(t);
		return int(textures.size()) - 1;
	}
};

namespace {

bool vec3_of(const JValue* v, float out[3], const char* what, std::string& err) {
	if (!v || v->kind != JValue::Array) { err = std::string("missing or non-array '") + what + "'"; return false; }
	out[0] = out[1] = out[2] = 0;  // SceneReader::getVec3 fills as many components as the array holds
	for (size_t i = 0; i < v->arr.size() && i < 3; i++) {
		if (v->arr[i].kind != JValue::Number) { err = std::string("non-numeric component in '") + what + "'"; return false; }
		out[i] = float(v->arr[i].num);
	}
	return true;
}
bool num_of(const JValue* v, float& out, const char* what, std::string& err) {
	if (!v || v->kind != JValue::Number) { err = std::string("missing or non-numeric '") + what + "'"; return false; }
	out = float(v->num);
	return true;
}
bool str_of(const JValue* v, std::string& out, const char* what, std::string& err) {
	if (!v || v->kind != JValue::String) { err = std::string("missing or non-string '") + what + "'"; return false; }
	out = v->str;
	return true;
}

// SceneReader::processMaterial, io/SceneReader.cpp:67-222
int process_material(ne_b200_scene_file& S, const JValue& m) {
	std::string err, name, type;
	if (!str_of(m.get("name"), name, "name", err)) { set_error("material: " + err); return NE_B200_ERR_INVALID; }
	if (!str_of(m.get("type"), type, "type", err)) { set_error("material " + name + ": " + err); return NE_B200_ERR_INVALID; }
	ne_b200_material o{};
	o.albedo_tex = o.roughness_tex = o.metallic_tex = o.normal_tex = -1;
	o.volume = -1;
	o.env_tex = -1;
	auto bad = [&](const std::string& e) { set_error("material " + name + ": " + e); return NE_B200_ERR_INVALID; };
	if (type == "microfacet") {
		o.type = NE_B200_MAT_MICROFACET;
		const JValue* al = m.get("albedo");
		if (!al) return bad("missing 'albedo'");
		float rough, metal;
		if (!num_of(m.get("roughness"), rough, "roughness", err) || !num_of(m.get("metallic"), metal, "metallic", err)) return bad(err);
		o.metallic_tex = S.addConstTexture(&metal, 1);  // texture order as SceneReader creates them :93-121
		if (al->kind == JValue::String) {
			auto it = S.imageByPath.find(al->str);
			if (it != S.imageByPath.end()) o.albedo_tex = it->second;
			else {
				std::string data, perr;
				if (!read_file(S.resources + al->str, data)) return bad("couldn't read the file at " + S.resources + al->str);
				int w = 0, h = 0;
				std::unique_ptr<std::vector<uint8_t>> px(new std::vector<uint8_t>());
				if (!png_decode(data, w, h, *px, perr)) return bad(al->str + ": " + perr);
				ne_b200_texture t{};
				t.width = w; t.height = h;
				t.format = NE_B200_TEX_RGBA8;
				t.wrap_u = t.wrap_v = NE_B200_WRAP_MIRROR;  // NE_TEX_SAMPLER_UVW_MIRROR, ResourceManager.cpp:306
				t.texels = px->data();
				S.bytes.push_back(std::move(px));
				S.textures.push_back(t);
				o.albedo_tex = int(S.textures.size()) - 1;
				S.imageByPath[al->str] = o.albedo_tex;
			}
		} else {
			float a[3];
			if (!vec3_of(al, a, "albedo", err)) return bad(err);
			o.albedo_tex = S.addConstTexture(a, 3);
		}
		o.roughness_tex = S.addConstTexture(&rough, 1);
		if (m.has("normalMap")) {
			float n[3];
			if (!vec3_of(m.get("normalMap"), n, "normalMap", err)) return bad(err);
			o.normal_tex = S.addConstTexture(n, 3);
			o.has_normal_flag = 1;  // NORMAL is the last texture added (Q24, Material.h:38-44)
		}
	} else if (type == "emitter") {
		o.type = NE_B200_MAT_EMITTER;
		if (!vec3_of(m.get("albedo"), o.li, "albedo", err)) return bad(err);
	} else if (type == "directionalLight") {
		o.type = NE_B200_MAT_DIRECTIONAL;
		float p[3];
		if (!vec3_of(m.get("albedo"), o.li, "albedo", err) || !vec3_of(m.get("position"), p, "position", err)) return bad(err);
		float len = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
		for (int k = 0; k < 3; k++) o.direction[k] = (0.0f - p[k]) / len;  // normalize(vec3(0) - position) :158
	} else if (type == "infiniteAreaLight") {
		o.type = NE_B200_MAT_INFINITE;
		std::string path, data, perr;
		if (!str_of(m.get("path"), path, "path", err)) return bad(err);
		auto it = S.imageByPath.find(path);
		if (it != S.imageByPath.end()) o.env_tex = it->second;
		else {
			if (!read_file(S.resources + path, data)) return bad("couldn't read the file at " + S.resources + path);
			int w = 0, h = 0;
			std::unique_ptr<std::vector<uint8_t>> px(new std::vector<uint8_t>());
			if (!png_decode(data, w, h, *px, perr)) return bad(path + ": " + perr);
			ne_b200_texture t{};
			t.width = w; t.height = h;
			t.format = NE_B200_TEX_RGBA8;
			t.wrap_u = t.wrap_v = NE_B200_WRAP_MIRROR;
			t.texels = px->data();
			S.bytes.push_back(std::move(px));
			S.textures.push_back(t);
			o.env_tex = int(S.textures.size()) - 1;
			S.imageByPath[path] = o.env_tex;
		}
	} else if (type == "volume") {
		o.type = NE_B200_MAT_VOLUME;
		std::string phase;
		if (!vec3_of(m.get("scattering"), o.scattering, "scattering", err) || !vec3_of(m.get("absorption"), o.absorption, "absorption", err) ||
		    !str_of(m.get("phaseFunction"), phase, "phaseFunction", err) || !num_of(m.get("density"), o.density_multiplier, "density", err))
			return bad(err);
		if (phase == "isotropic") o.phase = NE_B200_PHASE_ISOTROPIC;
		else if (phase == "hg" || phase == "henyey-greenstein") {
			o.phase = NE_B200_PHASE_HG;
			if (!num_of(m.get("g"), o.g, "g", err)) return bad(err);
		} else return bad("unknown phaseFunction '" + phase + "'");
		if (m.has("path")) {
			std::string path;
			if (!str_of(m.get("path"), path, "path", err)) return bad(err);
			if (path.find(".vdb") != std::string::npos)
				return bad(path + ": reading .vdb needs OpenVDB; hand its leaf bricks to ne_b200_volume (n_leaves) instead - see INTEGRATION.md");
			if (path.find(".vol") == std::string::npos) return bad(path + ": neither .vdb nor .vol");
			auto it = S.volumeByPath.find(path);
			if (it != S.volumeByPath.end()) o.volume = it->second;
			else {
				std::string text;
				if (!read_file(S.resources + path, text)) return bad("couldn't read the file at " + S.resources + path);
				int32_t dims[3];
				std::vector<float> grid;
				int rc = vol_parse(text, dims, &grid);
				if (rc) return rc;
				ne_b200_volume v{};
				v.width = dims[0]; v.height = dims[1]; v.depth = dims[2];
				v.dense = S.keepF(std::move(grid));
				S.volumes.push_back(v);
				o.volume = int(S.volumes.size()) - 1;
				S.volumeByPath[path] = o.volume;
			}
		}
	} else {
		return bad("invalid material type '" + type + "'");
	}
	// ResourceManager::replaceMaterial(name, ...): a later material with the same name replaces the earlier one
	S.materials.push_back(o);
	S.materialByName[name] = int(S.materials.size()) - 1;
	return NE_B200_OK;
}

// SceneReader::processPrimitives, io/SceneReader.cpp:224-648
int process_primitive(ne_b200_scene_file& S, const JValue& p) {
	std::string err, name, type;
	if (!str_of(p.get("name"), name, "name", err)) { set_error("primitive: " + err); return NE_B200_ERR_INVALID; }
	if (!str_of(p.get("type"), type, "type", err)) { set_error("primitive " + name + ": " + err); return NE_B200_ERR_INVALID; }
	auto bad = [&](const std::string& e) { set_error("primitive " + name + ": " + e); return NE_B200_ERR_INVALID; };
	const JValue* tr = p.get("transform");
	if (!tr) return bad("missing 'transform'");
	float pos[3], rot[3] = {0, 0, 0}, scale[3] = {1, 1, 1};
	if (!vec3_of(tr->get("position"), pos, "transform.position", err)) return bad(err);
	ne_b200_primitive o{};
	o.material = -1;
	o.collision = 1;
	if (const JValue* c = p.get("collision")) {
		if (c->kind != JValue::Bool) return bad("'collision' is not a bool");
		o.collision = c->b ? 1 : 0;
	}
	auto material = [&](bool required) -> int {
		const JValue* mn = p.get("materialName");
		if (!mn) return required ? -2 : -1;
		if (mn->kind != JValue::String) return -2;
		auto it = S.materialByName.find(mn->str);
		return it == S.materialByName.end() ? -2 : it->second;
	};
	bool needRS = type != "sphere";
	if (needRS && (!vec3_of(tr->get("scale"), scale, "transform.scale", err) || !vec3_of(tr->get("rotation"), rot, "transform.rotation", err))) return bad(err);
	if (type == "obj" || type == "gltf") {
		o.type = NE_B200_PRIM_MESH;
		o.material = material(false);
		if (o.material == -2) return bad("unknown materialName");
		std::string path, text;
		if (!str_of(p.get("path"), path, "path", err)) return bad(err);
		if (!read_file(S.resources + path, text)) return bad("couldn't read the file at " + S.resources + path);
		std::vector<float> vp, vuv;
		std::unique_ptr<std::vector<uint32_t>> idx(new std::vector<uint32_t>());
		bool hasUv = false;
		bool gltf = type == "gltf" || (text.size() >= 4 && !memcmp(text.data(), "glTF", 4));
		size_t slash = path.find_last_of('/');
		int rc = gltf ? gltf_parse(text, S.resources + (slash == std::string::npos ? "" : path.substr(0, slash + 1)), vp, vuv, *idx, hasUv)
		              : obj_parse(text, vp, vuv, *idx, hasUv);
		if (rc) return rc;
		o.n_vertices = int32_t(vp.size() / 3);
		o.n_triangles = int32_t(idx->size() / 3);
		o.positions = S.keepF(std::move(vp));
		o.uvs = hasUv ? S.keepF(std::move(vuv)) : nullptr;
		o.indices = idx->data();
		S.uints.push_back(std::move(idx));
	} else if (type == "point") {
		o.type = NE_B200_PRIM_POINT;
		for (int k = 0; k < 3; k++) o.point[k] = pos[k];  // the vertex AND the transform hold the position (Q25)
	} else if (type == "sphere") {
		o.type = NE_B200_PRIM_SPHERE;
		if (!num_of(p.get("radius"), o.radius, "radius", err)) return bad(err);
		// getTransform(pos, (0,0,0), (1,1,1)): scale and rotation of the JSON are ignored :376
	} else if (type == "rectangle") {
		o.type = NE_B200_PRIM_RECTANGLE;
	} else if (type == "volume") {
		o.type = NE_B200_PRIM_VOLUME;
	} else {
		return bad("invalid primitive type '" + type + "'");
	}
	if (o.type != NE_B200_PRIM_MESH) {
		o.material = material(true);
		if (o.material < 0) return bad("missing or unknown materialName");
	}
	ne_b200_make_transform(pos, rot, scale, o.to_world, o.to_object);
	S.primitives.push_back(o);
	return NE_B200_OK;
}

// SceneReader::processCameraAndRenderer, io/SceneReader.cpp:650-675
int process_camera(ne_b200_scene_file& S, const JValue* cam, const JValue* ren) {
	std::string err;
	auto bad = [&](const std::string& e) { set_error("camera/renderer: " + e); return NE_B200_ERR_INVALID; };
	if (!cam || !ren) return bad("missing 'camera' or 'renderer'");
	float up[3], speed, aperture;
	if (!vec3_of(cam->get("position"), S.camPosition, "camera.position", err) || !vec3_of(cam->get("lookAt"), S.camLookAt, "camera.lookAt", err) ||
	    !vec3_of(cam->get("up"), up, "camera.up", err) || !num_of(cam->get("speed"), speed, "camera.speed", err) ||
	    !num_of(cam->get("vfov"), S.vfov, "camera.vfov", err) || !num_of(cam->get("aperture"), aperture, "camera.aperture", err))
		return bad(err);
	const JValue* res = ren->get("resolution");
	if (!res || res->kind != JValue::Array || res->arr.size() < 2) return bad("missing renderer.resolution");
	S.settings.width = int32_t(res->arr[0].num);
	S.settings.height = int32_t(res->arr[1].num);
	const JValue* af = cam->get("autoFocus");
	if (af && af->kind == JValue::Bool && af->b) S.focus = 3.0f;  // (position - lookAt).length() is glm's component count (Q26)
	else if (!num_of(cam->get("focus"), S.focus, "camera.focus", err)) return bad(err);
	float spp, bounces;
	if (!num_of(ren->get("spp"), spp, "renderer.spp", err) || !num_of(ren->get("bounces"), bounces, "renderer.bounces", err)) return bad(err);
	S.settings.spp = int32_t(spp);
	S.settings.bounces = int32_t(bounces);
	const JValue* hdr = ren->get("HDR");
	if (!hdr || hdr->kind != JValue::Bool) return bad("missing renderer.HDR");
	S.settings.hdr = hdr->b ? 1 : 0;
	std::string mode;
	if (!str_of(ren->get("mode"), mode, "renderer.mode", err)) return bad(err);
	if (S.settings.width <= 0 || S.settings.height <= 0) return bad("bad resolution");
	return NE_B200_OK;
}

int build_scene(const std::string& text, const char* resources_dir, const std::string& label, ne_b200_scene_file** out) {
	JParser jp{text.c_str(), text.c_str() + text.size(), ""};
	JValue doc;
	if (!jp.parse(doc)) { set_error("malformatted json file: " + label + ": " + jp.err); return NE_B200_ERR_INVALID; }
	if (doc.kind != JValue::Object) { set_error("malformatted json file: " + label + ". Not a valid json object."); return NE_B200_ERR_INVALID; }
	if (!doc.has("version")) { set_error("incomplete json file: " + label + ". Version is not present."); return NE_B200_ERR_INVALID; }
	std::unique_ptr<ne_b200_scene_file> S(new ne_b200_scene_file());
	S->resources = resources_dir ? resources_dir : "";
	if (!S->resources.empty() && S->resources.back() != '/') S->resources += '/';
	const JValue* mats = doc.get("materials");
	const JValue* prims = doc.get("primitives");
	if (!mats || mats->kind != JValue::Array || !prims || prims->kind != JValue::Array) { set_error(label + ": missing 'materials' or 'primitives'"); return NE_B200_ERR_INVALID; }
	int rc;
	for (const JValue& m : mats->arr)
		if ((rc = process_material(*S, m))) return rc;
	for (const JValue& p : prims->arr)
		if ((rc = process_primitive(*S, p))) return rc;
	if ((rc = process_camera(*S, doc.get("camera"), doc.get("renderer")))) return rc;
	S->desc.n_textures = int32_t(S->textures.size()); S->desc.textures = S->textures.data();
	S->desc.n_volumes = int32_t(S->volumes.size()); S->desc.volumes = S->volumes.data();
	S->desc.n_materials = int32_t(S->materials.size()); S->desc.materials = S->materials.data();
	S->desc.n_primitives = int32_t(S->primitives.size()); S->desc.primitives = S->primitives.data();
	S->desc.sort_and_group = 0;
	*out = S.release();
	return NE_B200_OK;
}

}  // namespace

// No C++ exception may cross the C ABI (the caller is ctypes or the host engine: an unwinding std::bad_alloc would abort
// the process). Every entry point that parses untrusted files or allocates runs inside this barrier.
#define NE_GUARDED(...)                                                                                        \
	try {                                                                                                      \
		__VA_ARGS__                                                                                            \
	} catch (const std::bad_alloc&) {                                                                          \
		set_error("out of host memory");                                                                      \
		return NE_B200_ERR_NOMEM;                                                                              \
	} catch (const std::exception& e) {                                                                        \
		set_error(std::string("malformed input: ") + e.what());                                               \
		return NE_B200_ERR_INVALID;                                                                            \
	} catch (...) {                                                                                            \
		set_error("unknown failure");                                                                         \
		return NE_B200_ERR_INVALID;                                                                            \
	}

extern "C" {

int ne_b200_scene_file_load(const char* json_path, const char* resources_dir, ne_b200_scene_file** out) {
	NE_GUARDED(
	if (!json_path || !out) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	*out = nullptr;
	std::string text;
	if (!read_file(json_path, text)) { set_error(std::string("couldn't read ") + json_path); return NE_B200_ERR_INVALID; }
	return build_scene(text, resources_dir, json_path, out);
	)
}
int ne_b200_scene_file_parse(const char* json_text, const char* resources_dir, ne_b200_scene_file** out) {
	NE_GUARDED(
	if (!json_text || !out) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	*out = nullptr;
	return build_scene(json_text, resources_dir, "<text>", out);
	)
}
const ne_b200_scene_desc* ne_b200_scene_file_desc(const ne_b200_scene_file* f) { return f ? &f->desc : nullptr; }
int ne_b200_scene_file_camera(const ne_b200_scene_file* f, ne_b200_camera* out) {
	NE_GUARDED(
	if (!f || !out) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	const float up[3] = {0, 1, 0};  // the JSON's up and aperture are read and ignored (Q26, SceneReader.cpp:668)
	return ne_b200_camera_make(f->camPosition, f->camLookAt, up, f->vfov, float(f->settings.width) / float(f->settings.height), 0.0001f, f->focus, out);
	)
}
int ne_b200_scene_file_settings(const ne_b200_scene_file* f, ne_b200_render_settings* out) {
	NE_GUARDED(
	if (!f || !out) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	*out = f->settings;
	return NE_B200_OK;
	)
}
void ne_b200_scene_file_free(ne_b200_scene_file* f) { delete f; }

int ne_b200_vol_read(const char* path, int32_t dims[3], float* voxels) {
	NE_GUARDED(
	if (!path || !dims) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	std::string text;
	if (!read_file(path, text)) { set_error(std::string("couldn't read the file at ") + path); return NE_B200_ERR_INVALID; }
	if (!voxels) return vol_parse(text, dims, nullptr);
	std::vector<float> grid;
	int rc = vol_parse(text, dims, &grid);
	if (rc) return rc;
	memcpy(voxels, grid.data(), grid.size() * sizeof(float));
	return NE_B200_OK;
	)
}
int ne_b200_vol_write(const char* path, const int32_t dims[3], const float* voxels) {
	NE_GUARDED(
	if (!path || !dims || !voxels || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	FILE* f = fopen(path, "wb");
	if (!f) { set_error(std::string("couldn't write ") + path); return NE_B200_ERR_INVALID; }
	// every token is followed by ONE space, values on a single line (the reference's parser drops a value glued to a newline)
	fprintf(f, "%d %d %d \n", dims[0], dims[1], dims[2]);
	fprintf(f, "density\n");
	size_t n = size_t(dims[0]) * dims[1] * dims[2];
	std::string buf;
	buf.reserve(1 << 20);
	char tmp[32];
	for (size_t i = 0; i < n; i++) {
		int len = snprintf(tmp, sizeof(tmp), "%.9g ", voxels[i]);
		buf.append(tmp, len);
		if (buf.size() > (1 << 20) - 64) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
	}
	buf += "\n";
	fwrite(buf.data(), 1, buf.size(), f);
	fclose(f);
	return NE_B200_OK;
	)
}

int ne_b200_image_read_png(const char* path, int32_t dims[2], uint8_t* rgba) {
	NE_GUARDED(
	if (!path || !dims) { set_error("null argument"); return NE_B200_ERR_INVALID; }
	std::string data, err;
	if (!read_file(path, data)) { set_error(std::string("couldn't read the file at ") + path); return NE_B200_ERR_INVALID; }
	int w = 0, h = 0;
	std::vector<uint8_t> px;
	if (!png_decode(data, w, h, px, err)) { set_error(std::string(path) + ": " + err); return NE_B200_ERR_INVALID; }
	dims[0] = w; dims[1] = h;
	if (rgba) memcpy(rgba, px.data(), px.size());
	return NE_B200_OK;
	)
}

// saveImage(..., RGB32F, PNG, path), materials/Texture.h:48-62: clamp to [0,1], (uint8_t)(v * 255) truncation, 3 channels.
int ne_b200_image_write_png(const char* path, int width, int height, const float* rgb) {
	NE_GUARDED(
	if (!path || !rgb || width <= 0 || height <= 0) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	std::vector<uint8_t> raw((size_t(width) * 3 + 1) * height);
	for (int y = 0; y < height; y++) {
		uint8_t* row = &raw[(size_t(width) * 3 + 1) * y];
		row[0] = 0;
		for (int i = 0; i < width * 3; i++) {
			float v = rgb[size_t(y) * width * 3 + i];
			v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);  // NaN fails both tests and stays: (uint8_t)NaN is what the reference writes too
			row[1 + i] = uint8_t(v * 255);
		}
	}
	uLongf clen = compressBound(uLong(raw.size()));
	std::vector<uint8_t> comp(clen);
	if (compress2(comp.data(), &clen, raw.data(), uLong(raw.size()), 6) != Z_OK) { set_error("deflate failed"); return NE_B200_ERR_INVALID; }
	comp.resize(clen);
	std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	std::vector<uint8_t> ihdr;
	put_be32(ihdr, uint32_t(width));
	put_be32(ihdr, uint32_t(height));
	ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
	png_chunk(out, "IHDR", ihdr);
	png_chunk(out, "IDAT", comp);
	png_chunk(out, "IEND", {});
	FILE* f = fopen(path, "wb");
	if (!f) { set_error(std::string("couldn't write ") + path); return NE_B200_ERR_INVALID; }
	fwrite(out.data(), 1, out.size(), f);
	fclose(f);
	return NE_B200_OK;
	)
}

// saveImage(..., EXR, path) = tinyexr SaveEXR(data, w, h, 3, /*fp16*/0, path): a single-part scanline OpenEXR file with
// three FLOAT channels. Written uncompressed (tinyexr would zip it; the pixels any reader gets back are the same).
int ne_b200_image_write_exr(const char* path, int width, int height, const float* rgb) {
	NE_GUARDED(
	if (!path || !rgb || width <= 0 || height <= 0) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	std::vector<uint8_t> o;
	auto put32 = [&](uint32_t x) { for (int k = 0; k < 4; k++) o.push_back(uint8_t(x >> (8 * k))); };
	auto put64 = [&](uint64_t x) { for (int k = 0; k < 8; k++) o.push_back(uint8_t(x >> (8 * k))); };
	auto putS = [&](const char* s) { while (*s) o.push_back(uint8_t(*s++)); o.push_back(0); };
	auto putF = [&](float f) { uint32_t u; memcpy(&u, &f, 4); put32(u); };
	auto attr = [&](const char* name, const char* type, uint32_t size) { putS(name); putS(type); put32(size); };
	put32(20000630u);  // magic 0x76 0x2f 0x31 0x01
	put32(2u);         // version 2, single-part scanline
	attr("channels", "chlist", 3 * 18 + 1);
	for (const char* c : {"B", "G", "R"}) {  // alphabetical
		putS(c);
		put32(2u);  // FLOAT
		o.push_back(0); o.push_back(0); o.push_back(0); o.push_back(0);  // pLinear + reserved
		put32(1u); put32(1u);  // sampling
	}
	o.push_back(0);
	attr("compression", "compression", 1); o.push_back(0);
	attr("dataWindow", "box2i", 16); put32(0); put32(0); put32(uint32_t(width - 1)); put32(uint32_t(height - 1));
	attr("displayWindow", "box2i", 16); put32(0); put32(0); put32(uint32_t(width - 1)); put32(uint32_t(height - 1));
	attr("lineOrder", "lineOrder", 1); o.push_back(0);
	attr("pixelAspectRatio", "float", 4); putF(1.0f);
	attr("screenWindowCenter", "v2f", 8); putF(0.0f); putF(0.0f);
	attr("screenWindowWidth", "float", 4); putF(1.0f);
	o.push_back(0);  // end of header
	const uint64_t lineBytes = 8 + uint64_t(width) * 12;
	const uint64_t table = o.size() + uint64_t(height) * 8;
	for (int y = 0; y < height; y++) put64(table + lineBytes * y);
	for (int y = 0; y < height; y++) {
		put32(uint32_t(y));
		put32(uint32_t(width) * 12);
		for (int c = 2; c >= 0; c--)  // B, G, R planes
			for (int x = 0; x < width; x++) putF(rgb[(size_t(y) * width + x) * 3 + c]);
	}
	FILE* f = fopen(path, "wb");
	if (!f) { set_error(std::string("couldn't write ") + path); return NE_B200_ERR_INVALID; }
	fwrite(o.data(), 1, o.size(), f);
	fclose(f);
	return NE_B200_OK;
	)
}

// OfflineEngine::coreLoop's output.ppm (core/OfflineEngine.cpp:82,119-138): "P6\nW H\n65535\n", 16-bit big-endian
// samples, pixels written from the LAST to the first (so the file holds the frame rotated by 180 degrees), each channel
// uint16_t(value * 65535) of the tone-mapped pixel.
int ne_b200_image_write_ppm(const char* path, int width, int height, const float* rgb) {
	NE_GUARDED(
	if (!path || !rgb || width <= 0 || height <= 0) { set_error("bad argument"); return NE_B200_ERR_INVALID; }
	FILE* f = fopen(path, "wb");
	if (!f) { set_error(std::string("couldn't write ") + path); return NE_B200_ERR_INVALID; }
	fprintf(f, "P6\n%d %d\n%d\n", width, height, 65535);
	std::vector<uint8_t> o;
	o.reserve(size_t(width) * height * 6);
	for (long i = long(width) * height - 1; i >= 0; i--)
		for (int c = 0; c < 3; c++) {
			uint16_t v = uint16_t(rgb[size_t(i) * 3 + c] * 65535);
			o.push_back(uint8_t(v >> 8));
			o.push_back(uint8_t(v & 0xFF));
		}
	fwrite(o.data(), 1, o.size(), f);
	fclose(f);
	return NE_B200_OK;
	)
}

}  // extern "C"
