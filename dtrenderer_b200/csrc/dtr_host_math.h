// dtr_host_math.h -- the per-draw-call host arithmetic of the draw path.
//
// Everything that the reference evaluates ONCE per draw call stays on the host, because it uses
// libm cosf/sinf whose last-ulp behaviour differs between glibc and CUDA (SURVEY.md §7 hard part
// 4): the mesh matrix (DTRRender_Mesh, DTRendererRender.cpp:1404-1431), the 2D rotation/scale
// basis (TransformPoints :279-285) and the rectangle/bitmap corner transform
// (TransformRectPoints :378-393).  fp32, one rounding per operator, reference order; this
// translation unit is compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "dtr_records.h"

namespace dtr
{

struct Mat4
{
	float e[4][4]; // e[col][row] (dqn.h:849-855)
};

inline Mat4 mat4_identity()
{
	Mat4 m;
	std::memset(&m, 0, sizeof(m));
	m.e[0][0] = m.e[1][1] = m.e[2][2] = m.e[3][3] = 1.0f;
	return m;
}

// DqnMat4_Mul (dqn.h:2983-2997)
inline Mat4 mat4_mul(const Mat4 &a, const Mat4 &b)
{
	Mat4 r;
	for (int j = 0; j < 4; j++)
		for (int i = 0; i < 4; i++)
		{
			float s   = a.e[0][i] * b.e[j][0];
			s         = s + a.e[1][i] * b.e[j][1];
			s         = s + a.e[2][i] * b.e[j][2];
			s         = s + a.e[3][i] * b.e[j][3];
			r.e[j][i] = s;
		}
	return r;
}

struct Vec3
{
	float x, y, z;
};

inline Vec3 vec3_normalise(Vec3 a)
{
	float len = sqrtf(((a.x * a.x) + (a.y * a.y)) + (a.z * a.z));
	float inv = 1.0f / len;
	return Vec3{a.x * inv, a.y * inv, a.z * inv};
}
inline float vec3_dot(Vec3 a, Vec3 b)
{
	float r = 0.0f;
	r       = r + (a.x * b.x);
	r       = r + (a.y * b.y);
	r       = r + (a.z * b.z);
	return r;
}
inline Vec3 vec3_cross(Vec3 a, Vec3 b)
{
	return Vec3{(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}

// viewport * (perspective * (view * (T * (R * S)))) exactly as DTRRender_Mesh builds it.
inline Mat4 mesh_matrix(int width, int height, const float pos[3], float rotationDegrees,
                        const float axis[3], const float scale[3])
{
	Mat4 T    = mat4_identity();
	T.e[3][0] = pos[0];
	T.e[3][1] = pos[1];
	T.e[3][2] = pos[2];
	Mat4 S;
	std::memset(&S, 0, sizeof(S));
	S.e[0][0] = scale[0];
	S.e[1][1] = scale[1];
	S.e[2][2] = scale[2];
	S.e[3][3] = 1.0f;

	// DQN_DEGREES_TO_RADIANS with DQN_PI = 3.14159265359f (dqn.h:123,126); DqnMat4_Rotate does
	// not normalise its axis (dqn.h:2941-2961)
	float radians = rotationDegrees * (3.14159265359f / 180.0f);
	float x = axis[0], y = axis[1], z = axis[2];
	float sv = sinf(radians), cv = cosf(radians), omc = 1.0f - cv;
	Mat4  R   = mat4_identity();
	R.e[0][0] = ((x * x) * omc) + cv;
	R.e[0][1] = ((x * y) * omc) + (z * sv);
	R.e[0][2] = ((x * z) * omc) - (y * sv);
	R.e[1][0] = ((y * x) * omc) - (z * sv);
	R.e[1][1] = ((y * y) * omc) + cv;
	R.e[1][2] = ((y * z) * omc) + (x * sv);
	R.e[2][0] = ((z * x) * omc) + (y * sv);
	R.e[2][1] = ((z * y) * omc) - (x * sv);
	R.e[2][2] = ((z * z) * omc) + cv;
	Mat4 model = mat4_mul(T, mat4_mul(R, S));

	// DqnMat4_LookAt(eye (0,0,1), center 0, up +Y) (dqn.h:2905-2930)
	Vec3 eye = {0, 0, 1}, up = {0, 1, 0}, center = {0, 0, 0};
	Vec3 f   = vec3_normalise(Vec3{eye.x - center.x, eye.y - center.y, eye.z - center.z});
	Vec3 s   = vec3_normalise(vec3_cross(up, f));
	Vec3 u   = vec3_cross(f, s);
	Mat4 V;
	std::memset(&V, 0, sizeof(V));
	V.e[0][0] = s.x; V.e[0][1] = u.x; V.e[0][2] = f.x;
	V.e[1][0] = s.y; V.e[1][1] = u.y; V.e[1][2] = f.y;
	V.e[2][0] = s.z; V.e[2][1] = u.z; V.e[2][2] = f.z;
	V.e[3][0] = vec3_dot(s, eye);
	V.e[3][1] = vec3_dot(u, eye);
	V.e[3][2] = -vec3_dot(f, eye);
	V.e[3][3] = 1.0f;

	// the reference discards DqnMat4_Perspective and uses identity with e[2][3] = -1/|eye-center|
	Mat4  Pm = mat4_identity();
	float dx = center.x - eye.x, dy = center.y - eye.y, dz = center.z - eye.z;
	float lensq = ((dx * dx) + (dy * dy)) + (dz * dz);
	float len   = (lensq == 0) ? 0.0f : sqrtf(lensq);
	Pm.e[2][3]  = -1.0f / len;

	// GLViewport(0, 0, W, H): z maps to [0, 255] (:1238-1263)
	Mat4  VP = mat4_identity();
	float hw = (float)width * 0.5f, hh = (float)height * 0.5f, hd = 255.0f * 0.5f;
	VP.e[0][0] = hw;
	VP.e[1][1] = hh;
	VP.e[2][2] = hd;
	VP.e[3][0] = 0.0f + hw;
	VP.e[3][1] = 0.0f + hh;
	VP.e[3][2] = hd;
	return mat4_mul(VP, mat4_mul(Pm, mat4_mul(V, model)));
}

struct Basis2
{
	float xAxis[2], yAxis[2];
};

// TransformPoints' axes (DTRendererRender.cpp:282-285): rotation in RADIANS
inline Basis2 make_basis(float rotation, float sx, float sy)
{
	Basis2 b;
	b.xAxis[0] = cosf(rotation);
	b.xAxis[1] = sinf(rotation);
	b.yAxis[0] = -b.xAxis[1];
	b.yAxis[1] = b.xAxis[0];
	b.xAxis[0] = b.xAxis[0] * sx;
	b.xAxis[1] = b.xAxis[1] * sx;
	b.yAxis[0] = b.yAxis[0] * sy;
	b.yAxis[1] = b.yAxis[1] * sy;
	return b;
}

inline float ref_maxf(float a, float b) { return (a < b) ? b : a; } // DQN_MAX
inline float ref_minf(float a, float b) { return (a < b) ? a : b; } // DQN_MIN

// number of integers i >= 0 with (float)i < s  (the reference's `for (i32 i = 0; i < s; i++)`)
inline int loop_count(float s)
{
	if (!(s > 0.0f)) return 0;
	if (s > 65536.0f) return 65536;
	return (int)ceilf(s);
}

inline uint32_t f2u(float v)
{
	uint32_t u;
	std::memcpy(&u, &v, 4);
	return u;
}

inline void to_linear_premul(const float c[4], float out[4])
{
	out[0] = c[0] * c[0]; // DTRRender_SRGB1ToLinearSpaceV4 then PreMultiplyAlpha1 (:36-53,76-92)
	out[1] = c[1] * c[1];
	out[2] = c[2] * c[2];
	out[3] = c[3];
	out[0] = out[0] * out[3];
	out[1] = out[1] * out[3];
	out[2] = out[2] * out[3];
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Rectangle (texW == 0) or bitmap quad -> PrimRecord.  Returns false when nothing can be drawn.
// outPoints (optional): the four transformed corners, Basis/XAxis/Point/YAxis order;
// outBounds (optional): their unclipped bounding box {minx, miny, maxx, maxy} -- both are filled
// even when the quad itself is invisible (the reference's debug overlay still draws then).
inline bool setup_quad(int W, int H, const float mn[2], const float mx[2], float rotation,
                       const float anchor[2], const float scale[2], const float color[4], bool bitmap,
                       int texId, int texW, int texH, PrimRecord *rec, float (*outPoints)[2] = nullptr,
                       float *outBounds = nullptr)
{
	std::memset(rec, 0, sizeof(*rec));
	// TransformRectPoints (:378-393)
	float  dimw = mx[0] - mn[0], dimh = mx[1] - mn[1];
	float  ox = mn[0] + (anchor[0] * dimw), oy = mn[1] + (anchor[1] * dimh);
	float  in[4][2] = {{mn[0] - ox, mn[1] - oy}, {mx[0] - ox, mn[1] - oy}, {mx[0] - ox, mx[1] - oy}, {mn[0] - ox, mx[1] - oy}};
	Basis2 bs = make_basis(rotation, scale[0], scale[1]);
	float  p[4][2];
	for (int i = 0; i < 4; i++)
	{
		p[i][0] = (ox + (bs.xAxis[0] * in[i][0])) + (bs.yAxis[0] * in[i][1]);
		p[i][1] = (oy + (bs.xAxis[1] * in[i][0])) + (bs.yAxis[1] * in[i][1]);
	}
	float bminx = p[0][0], bminy = p[0][1], bmaxx = p[0][0], bmaxy = p[0][1];
	for (int i = 1; i < 4; i++)
	{
		bminx = ref_minf(bminx, p[i][0]); bminy = ref_minf(bminy, p[i][1]);
		bmaxx = ref_maxf(bmaxx, p[i][0]); bmaxy = ref_maxf(bmaxy, p[i][1]);
	}
	if (outPoints)
		for (int i = 0; i < 4; i++)
		{
			outPoints[i][0] = p[i][0];
			outPoints[i][1] = p[i][1];
		}
	if (outBounds)
	{
		outBounds[0] = bminx; outBounds[1] = bminy; outBounds[2] = bmaxx; outBounds[3] = bmaxy;
	}
	// clip to (0,0)-(W,H) (:436-440,1628-1632; dqn.h:3071-3081)
	float cmaxx = ref_minf(bmaxx, (float)W - 0.0f), cmaxy = ref_minf(bmaxy, (float)H - 0.0f);
	float cminx = ref_maxf(0.0f, bminx), cminy = ref_maxf(0.0f, bminy);
	float sizew = cmaxx - cminx, sizeh = cmaxy - cminy;

	int nx, ny;
	uint32_t type;
	if (bitmap)
	{
		type = PRIM_BITMAP;
		nx   = (sizew > 0.0f) ? (int)sizew : 0; // `x < (i32)clippedSize.w` (:1646-1649)
		ny   = (sizeh > 0.0f) ? (int)sizeh : 0;
	}
	else if (rotation != 0)
	{
		type = PRIM_RECT_ROT; // the rotated loop swaps its w/h bounds (:450-453)
		ny   = loop_count(sizew);
		nx   = loop_count(sizeh);
	}
	else
	{
		type = PRIM_RECT_FILL;
		ny   = loop_count(sizeh);
		nx   = loop_count(sizew);
	}
	if (!(cminx < 65536.0f) || !(cminy < 65536.0f)) return false;
	int sx = (int)cminx, sy = (int)cminy;
	// SetPixel rejects pixels outside the buffer (:129-130)
	int x0 = clampi(sx, 0, W), y0 = clampi(sy, 0, H), x1 = clampi(sx + nx, 0, W), y1 = clampi(sy + ny, 0, H);
	if (x1 <= x0 || y1 <= y0) return false;

	float col[4];
	to_linear_premul(color, col);
	rec->w[QW_FLAGS] = type;
	rec->w[QW_TEX]   = (uint32_t)texId;
	rec->w[QW_MIN]   = (uint32_t)x0 | ((uint32_t)y0 << 16);
	rec->w[QW_MAX]   = (uint32_t)x1 | ((uint32_t)y1 << 16);
	for (int i = 0; i < 4; i++)
	{
		rec->w[QW_P + 2 * i + 0] = f2u(p[i][0]);
		rec->w[QW_P + 2 * i + 1] = f2u(p[i][1]);
	}
	for (int i = 0; i < 4; i++) rec->w[QW_COLOR + i] = f2u(col[i]);
	if (bitmap)
	{
		// 1 / |axis|^2 with DqnV2_LengthSquared(0, axis) (:1644-1645; dqn.h:2439-2445)
		float xa0 = p[1][0] - p[0][0], xa1 = p[1][1] - p[0][1];
		float ya0 = p[3][0] - p[0][0], ya1 = p[3][1] - p[0][1];
		float tx = xa0 - 0.0f, ty = xa1 - 0.0f;
		rec->w[QW_INVX] = f2u(1.0f / ((tx * tx) + (ty * ty)));
		tx = ya0 - 0.0f; ty = ya1 - 0.0f;
		rec->w[QW_INVY]   = f2u(1.0f / ((tx * tx) + (ty * ty)));
		rec->w[QW_TEXDIM] = (uint32_t)texW | ((uint32_t)texH << 16);
	}
	return true;
}

// DTRRender_Clear's packed colour (:1801-1811): truncating conversion
inline uint32_t pack_clear(const float rgb[3])
{
	float r = rgb[0] * 255.0f, g = rgb[1] * 255.0f, b = rgb[2] * 255.0f;
	return (uint32_t)(((int32_t)0 << 24) | ((int32_t)r << 16) | ((int32_t)g << 8) | ((int32_t)b << 0));
}

} // namespace dtr
