/* dtr_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/dtro.h).
 *
 * Plain-C restatement of the DTRenderer scalar draw path: the CPU checker the CUDA back end is
 * compared against on machines where /root/reference does not exist (the GPU box).  It is NOT
 * shipped and NOT on the product path: dtrenderer_b200/ never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pinning.py compares this file, bit for bit (colour
 * and z), with the unmodified reference compiled by oracle/Makefile (`make ref`) on every scene
 * family, and tests/golden/ holds frame digests produced by that reference build
 * (tests/golden/make_golden.py).
 *
 * Arithmetic contract (SURVEY.md §8a'): IEEE fp32, one rounding per operator, evaluated in the
 * reference's order; build with -ffp-contract=off and never -ffast-math.  Each function cites
 * the reference lines it restates (paths relative to /root/reference/src).
 *
 * Structure differs from the reference on purpose: every draw call is split into a per-primitive
 * SETUP (what the CUDA setup kernel computes once) and a per-pixel RASTER/SHADE step (what the
 * CUDA tile kernel evaluates), so the same decomposition can be read in both places.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dtro.h"

struct dtro_ctx
{
	int       width, height;
	uint32_t *color;
	float    *z;
	uint64_t  counters[2];
};

typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;
typedef struct { float r, g, b, a; } rgba;
typedef struct { float e[4][4]; } mat4; /* e[col][row], dqn.h:849-855 */

/* dqn.h:129-130 -- note the comparison direction, it decides what NaNs and +-0 do */
#define ORA_MAX(a, b) ((a) < (b) ? (b) : (a))
#define ORA_MIN(a, b) ((a) < (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------ */
/* colour pipeline: DTRendererRender.cpp:36-53,76-111,124-191                                  */
/* ------------------------------------------------------------------------------------------ */

/* sRGB->linear is x*x, then premultiply rgb by a (DTRRender_SRGB1ToLinearSpaceV4 + PreMultiplyAlpha1) */
static rgba to_linear_premul(const float c[4])
{
	rgba o;
	o.r = c[0] * c[0];
	o.g = c[1] * c[1];
	o.b = c[2] * c[2];
	o.a = c[3];
	o.r = o.r * o.a;
	o.g = o.g * o.a;
	o.b = o.b * o.a;
	return o;
}

static float dst_channel_linear(uint32_t byte)
{
	/* DTRRENDER_INV_255 is the unparenthesised `1.0f/255.0f` (DTRendererRender.h:7), so
	 * `(f32)b * DTRRENDER_INV_255` parses as ((f32)b * 1.0f) / 255.0f: a true division. */
	float s = ((float)byte * 1.0f) / 255.0f;
	return s * s;
}

static float out_channel(float src, float inv_a, float dst_lin)
{
	float d = src + (inv_a * dst_lin);
	d       = (d == 0) ? 0.0f : sqrtf(d); /* DTRRender_LinearToSRGB1Spacef :94-100 */
	d       = d * 255.0f;
	if (d > 255.0f) d = 255.0f;
	return d;
}

/* SetPixel with ColorSpace_Linear (:124-191): the only colour space the hot path uses */
static void blend_pixel(dtro_ctx *c, int x, int y, rgba src)
{
	if (x < 0 || x > c->width - 1) return;
	if (y < 0 || y > c->height - 1) return;
	uint32_t *px  = &c->color[(size_t)x + (size_t)y * (size_t)c->width];
	uint32_t  dst = *px;
	float     dr  = dst_channel_linear((dst >> 16) & 0xFF);
	float     dg  = dst_channel_linear((dst >> 8) & 0xFF);
	float     db  = dst_channel_linear((dst >> 0) & 0xFF);
	float     inv = 1 - src.a;
	float     r   = out_channel(src.r, inv, dr);
	float     g   = out_channel(src.g, inv, dg);
	float     b   = out_channel(src.b, inv, db);
	*px           = ((uint32_t)r << 16) | ((uint32_t)g << 8) | ((uint32_t)b << 0);
	c->counters[DTRO_COUNTER_SETPIXELS]++;
}

/* ------------------------------------------------------------------------------------------ */
/* 2D transform helpers: DTRendererRender.cpp:275-292,378-413,605-617; dqn.h:3071-3081        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { v2 xAxis, yAxis; } basis2;

static basis2 make_basis(float rotation, float sx, float sy)
{
	basis2 b;
	b.xAxis.x = cosf(rotation);
	b.xAxis.y = sinf(rotation);
	b.yAxis.x = -b.xAxis.y;
	b.yAxis.y = b.xAxis.x;
	b.xAxis.x = b.xAxis.x * sx;
	b.xAxis.y = b.xAxis.y * sx;
	b.yAxis.x = b.yAxis.x * sy;
	b.yAxis.y = b.yAxis.y * sy;
	return b;
}

/* origin + xAxis*p.x + yAxis*p.y, left to right (:287-291) */
static v2 apply_basis(v2 origin, basis2 b, v2 p)
{
	v2 o;
	o.x = (origin.x + (b.xAxis.x * p.x)) + (b.yAxis.x * p.y);
	o.y = (origin.y + (b.xAxis.y * p.x)) + (b.yAxis.y * p.y);
	return o;
}

/* ------------------------------------------------------------------------------------------ */
/* triangle: setup = TexturedTriangleInternal (:1265-1350) + SlowTriangle preamble (:1104-1145) */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
	int   minx, miny, maxx, maxy;    /* loop bounds [min, max) */
	float e0[3], dx[3], dy[3];       /* edge functions at (minx,miny) and their steps */
	float inv_area;
	float z1, dz2, dz3;
	rgba  color;                     /* linear, premultiplied */
	float light[3][3];               /* [vertex][rgb] = color.rgb * max(0, I_k) */
	int   ignore_light;
	v2    uv1, duv2, duv3;
	int   skip;                      /* area == 0 */
} tri_setup;

static float edge_fn(v2 a, v2 b, v2 c)
{
	return ((b.x - a.x) * (c.y - a.y)) - ((b.y - a.y) * (c.x - a.x)); /* :532-536 */
}

static v3 v3_normalise(v3 a)
{
	float len = sqrtf(((a.x * a.x) + (a.y * a.y)) + (a.z * a.z)); /* dqn.h:2716-2723 */
	float inv = 1 / len;
	v3 o      = {a.x * inv, a.y * inv, a.z * inv};
	return o;
}

static float v3_dot(v3 a, v3 b)
{
	float r = 0; /* dqn.h:2676-2690: accumulates from 0, left to right */
	r += (a.x * b.x);
	r += (a.y * b.y);
	r += (a.z * b.z);
	return r;
}

static v3 v3_cross(v3 a, v3 b)
{
	v3 o;
	o.x = (a.y * b.z) - (a.z * b.y);
	o.y = (a.z * b.x) - (a.x * b.z);
	o.z = (a.x * b.y) - (a.y * b.x);
	return o;
}

static void setup_triangle(const dtro_ctx *c, v3 p1, v3 p2, v3 p3, v2 uv1, v2 uv2, v2 uv3,
                           int light_mode, v3 light_vec, const v3 normals[3],
                           const float color_in[4], const float transform[7], tri_setup *s)
{
	/* winding: swaps the POSITIONS of p2/p3 only, uv and normals stay (:24-34,1276) */
	float area2 = (((p2.x - p1.x) * (p2.y + p1.y)) + ((p3.x - p2.x) * (p3.y + p2.y))) +
	              ((p1.x - p3.x) * (p1.y + p3.y));
	if (area2 > 0)
	{
		v3 t = p2;
		p2   = p3;
		p3   = t;
	}

	/* anchor origin (:605-617) and the round trip through it (:1278-1284) */
	float ax = transform[1], ay = transform[2];
	v2 origin;
	origin.x = (p1.x + ((p2.x - p1.x) * ax)) + ((p3.x - p1.x) * ax);
	origin.y = (p1.y + ((p2.y - p1.y) * ay)) + ((p3.y - p1.y) * ay);
	basis2 bs = make_basis(transform[0], transform[4], transform[5]);
	v2 q1 = {p1.x - origin.x, p1.y - origin.y};
	v2 q2 = {p2.x - origin.x, p2.y - origin.y};
	v2 q3 = {p3.x - origin.x, p3.y - origin.y};
	q1 = apply_basis(origin, bs, q1);
	q2 = apply_basis(origin, bs, q2);
	q3 = apply_basis(origin, bs, q3);
	p1.x = q1.x; p1.y = q1.y;
	p2.x = q2.x; p2.y = q2.y;
	p3.x = q3.x; p3.y = q3.y;

	/* bbox, clip to (0,0)-(W-1,H-1), truncate (:395-413,1286-1290; dqn.h:3071-3081) */
	float bminx = q1.x, bminy = q1.y, bmaxx = q1.x, bmaxy = q1.y;
	bminx = ORA_MIN(bminx, q2.x); bminy = ORA_MIN(bminy, q2.y);
	bmaxx = ORA_MAX(bmaxx, q2.x); bmaxy = ORA_MAX(bmaxy, q2.y);
	bminx = ORA_MIN(bminx, q3.x); bminy = ORA_MIN(bminy, q3.y);
	bmaxx = ORA_MAX(bmaxx, q3.x); bmaxy = ORA_MAX(bmaxy, q3.y);
	float clipw = (float)(c->width - 1) - (float)0, cliph = (float)(c->height - 1) - (float)0;
	bmaxx = ORA_MIN(bmaxx, clipw);
	bmaxy = ORA_MIN(bmaxy, cliph);
	bminx = ORA_MAX(0.0f, bminx);
	bminy = ORA_MAX(0.0f, bminy);
	s->minx = (int)bminx; s->miny = (int)bminy;
	s->maxx = (int)bmaxx; s->maxy = (int)bmaxy;

	/* lighting (:1295-1322) */
	float col[4] = {color_in[0], color_in[1], color_in[2], color_in[3]};
	float I[3]   = {1, 1, 1};
	s->ignore_light = 0;
	if (light_mode == DTRO_SHADE_FULLBRIGHT)
	{
		s->ignore_light = 1;
	}
	else
	{
		v3 L = v3_normalise(light_vec);
		if (light_mode == DTRO_SHADE_FLAT)
		{
			v3 a = {p2.x - p1.x, p2.y - p1.y, p2.z - p1.z};
			v3 b = {p3.x - p1.x, p3.y - p1.y, p3.z - p1.z};
			float intensity = v3_dot(v3_normalise(v3_cross(a, b)), L);
			intensity       = ORA_MAX(0, intensity);
			col[0] *= intensity; /* scales the sRGB colour BEFORE the square below */
			col[1] *= intensity;
			col[2] *= intensity;
		}
		else
		{
			I[0] = v3_dot(v3_normalise(normals[0]), L);
			I[1] = v3_dot(v3_normalise(normals[1]), L);
			I[2] = v3_dot(v3_normalise(normals[2]), L);
		}
	}

	/* SlowTriangle preamble (:1104-1145) */
	s->color = to_linear_premul(col);
	v2 start = {(float)s->minx, (float)s->miny};
	v2 a1 = {p1.x, p1.y}, a2 = {p2.x, p2.y}, a3 = {p3.x, p3.y};
	s->e0[0] = edge_fn(a2, a3, start); s->dx[0] = p2.y - p3.y; s->dy[0] = p3.x - p2.x;
	s->e0[1] = edge_fn(a3, a1, start); s->dx[1] = p3.y - p1.y; s->dy[1] = p1.x - p3.x;
	s->e0[2] = edge_fn(a1, a2, start); s->dx[2] = p1.y - p2.y; s->dy[2] = p2.x - p1.x;
	float area = (s->e0[0] + s->e0[1]) + s->e0[2];
	s->skip    = (area == 0);
	s->inv_area = 1.0f / area;
	s->z1  = p1.z;
	s->dz2 = p2.z - p1.z;
	s->dz3 = p3.z - p1.z;
	s->uv1 = uv1;
	s->duv2.x = uv2.x - uv1.x; s->duv2.y = uv2.y - uv1.y;
	s->duv3.x = uv3.x - uv1.x; s->duv3.y = uv3.y - uv1.y;
	for (int k = 0; k < 3; k++)
	{
		float m        = ORA_MAX(0, I[k]);
		s->light[k][0] = s->color.r * m;
		s->light[k][1] = s->color.g * m;
		s->light[k][2] = s->color.b * m;
	}
}

static float clampf(float v, float lo, float hi)
{
	if (v < lo) return lo; /* dqn.h:2325-2330 */
	if (v > hi) return hi;
	return v;
}

/* per covered pixel: SlowTriangle :1154-1222 */
static void shade_fragment(dtro_ctx *c, const tri_setup *s, const uint8_t *tex, int texW, int texH,
                           int x, int y, float e1, float e2, float e3)
{
	float bA = e1 * s->inv_area;
	float bB = e2 * s->inv_area;
	float bC = e3 * s->inv_area;
	size_t idx = (size_t)x + (size_t)y * (size_t)c->width;
	float z = (s->z1 + (bB * s->dz2)) + (bC * s->dz3);
	if (!(z > c->z[idx])) return;
	c->z[idx] = z; /* written even when the fragment is translucent */

	rgba f = s->color;
	if (!s->ignore_light)
	{
		float lr = ((s->light[0][0] * bA) + (s->light[1][0] * bB)) + (s->light[2][0] * bC);
		float lg = ((s->light[0][1] * bA) + (s->light[1][1] * bB)) + (s->light[2][1] * bC);
		float lb = ((s->light[0][2] * bA) + (s->light[1][2] * bB)) + (s->light[2][2] * bC);
		f.r = f.r * lr;
		f.g = f.g * lg;
		f.b = f.b * lb;
	}
	if (tex)
	{
		const float INV_255 = 1 / 255.0f; /* reciprocal multiply here (:1139,1211) */
		float u = (s->uv1.x + (s->duv2.x * bB)) + (s->duv3.x * bC);
		float v = (s->uv1.y + (s->duv2.y * bB)) + (s->duv3.y * bC);
		u = clampf(u, 0.0f, 1.0f);
		v = clampf(v, 0.0f, 1.0f);
		int tx = (int)(u * (float)texW); /* NEAREST (:1196-1203) */
		int ty = (int)(v * (float)texH);
		uint32_t t;
		memcpy(&t, tex + ((size_t)tx * 4 + (size_t)ty * (size_t)texW * 4), 4);
		float ta = (float)(t >> 24) * INV_255;
		float tb = (float)((t >> 16) & 0xFF) * INV_255;
		float tg = (float)((t >> 8) & 0xFF) * INV_255;
		float tr = (float)((t >> 0) & 0xFF) * INV_255;
		tr = tr * tr; tg = tg * tg; tb = tb * tb; /* alpha is not squared */
		f.r = f.r * tr; f.g = f.g * tg; f.b = f.b * tb; f.a = f.a * ta;
	}
	blend_pixel(c, x, y, f);
}

static void raster_triangle(dtro_ctx *c, const tri_setup *s, const uint8_t *tex, int texW, int texH)
{
	if (s->skip) return;
	float r1 = s->e0[0], r2 = s->e0[1], r3 = s->e0[2];
	for (int y = s->miny; y < s->maxy; y++)
	{
		float e1 = r1, e2 = r2, e3 = r3;
		for (int x = s->minx; x < s->maxx; x++)
		{
			if (e1 >= 0 && e2 >= 0 && e3 >= 0) shade_fragment(c, s, tex, texW, texH, x, y, e1, e2, e3);
			e1 += s->dx[0]; /* sequential fp32 accumulation is part of the contract (:1225-1232) */
			e2 += s->dx[1];
			e3 += s->dx[2];
		}
		r1 += s->dy[0];
		r2 += s->dy[1];
		r3 += s->dy[2];
	}
}

static void draw_triangle(dtro_ctx *c, const float p[9], const float uv[6], const uint8_t *tex,
                          int texW, int texH, int light_mode, const float light_vec[3],
                          const v3 normals[3], const float color[4], const float transform[7])
{
	tri_setup s;
	v3 p1 = {p[0], p[1], p[2]}, p2 = {p[3], p[4], p[5]}, p3 = {p[6], p[7], p[8]};
	v2 uv1 = {uv[0], uv[1]}, uv2 = {uv[2], uv[3]}, uv3 = {uv[4], uv[5]};
	v3 L = {0, 0, 0};
	if (light_vec) { L.x = light_vec[0]; L.y = light_vec[1]; L.z = light_vec[2]; }
	setup_triangle(c, p1, p2, p3, uv1, uv2, uv3, light_mode, L, normals, color, transform, &s);
	raster_triangle(c, &s, tex, texW, texH);
	c->counters[DTRO_COUNTER_TRIANGLES]++;
}

/* ------------------------------------------------------------------------------------------ */
/* mesh: DTRRender_Mesh (:1395-1585), matrices dqn.h:2905-3008, GLViewport (:1238-1263)        */
/* ------------------------------------------------------------------------------------------ */
static mat4 mat4_identity(void)
{
	mat4 m;
	memset(&m, 0, sizeof(m));
	m.e[0][0] = m.e[1][1] = m.e[2][2] = m.e[3][3] = 1;
	return m;
}

static mat4 mat4_mul(mat4 a, mat4 b)
{
	mat4 r;
	for (int j = 0; j < 4; j++)
		for (int i = 0; i < 4; i++)
			r.e[j][i] = ((a.e[0][i] * b.e[j][0] + a.e[1][i] * b.e[j][1]) + a.e[2][i] * b.e[j][2]) +
			            a.e[3][i] * b.e[j][3];
	return r;
}

static void mat4_mulv4(const mat4 *a, const float b[4], float out[4])
{
	for (int r = 0; r < 4; r++)
		out[r] = (((a->e[0][r] * b[0]) + (a->e[1][r] * b[1])) + (a->e[2][r] * b[2])) + (a->e[3][r] * b[3]);
}

static mat4 mesh_matrix(int width, int height, const float pos[3], const float transform[7])
{
	/* model = T * (R * S) (:1407-1412) */
	mat4 T = mat4_identity();
	T.e[3][0] = pos[0]; T.e[3][1] = pos[1]; T.e[3][2] = pos[2];
	mat4 S;
	memset(&S, 0, sizeof(S));
	S.e[0][0] = transform[4]; S.e[1][1] = transform[5]; S.e[2][2] = transform[6]; S.e[3][3] = 1;
	float radians = (transform[0] * (3.14159265359f / 180.0f)); /* dqn.h:123,126 */
	float x = transform[1], y = transform[2], z = transform[3]; /* axis = anchor, not normalised */
	float sv = sinf(radians), cv = cosf(radians), omc = 1 - cv;
	mat4 R = mat4_identity(); /* dqn.h:2941-2961 */
	R.e[0][0] = ((x * x) * omc) + cv;
	R.e[0][1] = (x * y * omc) + (z * sv);
	R.e[0][2] = (x * z * omc) - (y * sv);
	R.e[1][0] = (y * x * omc) - (z * sv);
	R.e[1][1] = ((y * y) * omc) + cv;
	R.e[1][2] = (y * z * omc) + (x * sv);
	R.e[2][0] = (z * x * omc) + (y * sv);
	R.e[2][1] = (z * y * omc) - (x * sv);
	R.e[2][2] = ((z * z) * omc) + cv;
	mat4 model = mat4_mul(T, mat4_mul(R, S));

	/* view = LookAt(eye (0,0,1), center 0, up +Y) (:1415-1418; dqn.h:2905-2930) */
	v3 eye = {0, 0, 1}, up = {0, 1, 0}, center = {0, 0, 0};
	v3 f = {eye.x - center.x, eye.y - center.y, eye.z - center.z};
	f    = v3_normalise(f);
	v3 s = v3_normalise(v3_cross(up, f));
	v3 u = v3_cross(f, s);
	mat4 V;
	memset(&V, 0, sizeof(V));
	V.e[0][0] = s.x; V.e[0][1] = u.x; V.e[0][2] = f.x;
	V.e[1][0] = s.y; V.e[1][1] = u.y; V.e[1][2] = f.y;
	V.e[2][0] = s.z; V.e[2][1] = u.z; V.e[2][2] = f.z;
	V.e[3][0] = v3_dot(s, eye);
	V.e[3][1] = v3_dot(u, eye);
	V.e[3][2] = -v3_dot(f, eye);
	V.e[3][3] = 1.0f;

	/* "perspective": identity with e[2][3] = -1/|eye-center| (:1421-1424) */
	mat4 P = mat4_identity();
	float dx = center.x - eye.x, dy = center.y - eye.y, dz = center.z - eye.z;
	float lensq = ((dx * dx) + (dy * dy)) + (dz * dz);
	float len   = (lensq == 0) ? 0 : sqrtf(lensq);
	P.e[2][3]   = -1.0f / len;

	/* viewport (:1238-1263) */
	mat4 VP = mat4_identity();
	float hw = (float)width * 0.5f, hh = (float)height * 0.5f, hd = 255.0f * 0.5f;
	VP.e[0][0] = hw; VP.e[1][1] = hh; VP.e[2][2] = hd;
	VP.e[3][0] = 0 + hw; VP.e[3][1] = 0 + hh; VP.e[3][2] = hd;

	return mat4_mul(VP, mat4_mul(P, mat4_mul(V, model))); /* :1428-1430 */
}

/* ------------------------------------------------------------------------------------------ */
/* rectangle / bitmap: DTRendererRender.cpp:378-513,1596-1791                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
	v2    p[4];                       /* Basis, XAxis, Point, YAxis */
	float cminx, cminy, csizew, csizeh; /* clipped draw rect: min and size (floats) */
	float bminx, bminy, bmaxx, bmaxy; /* unclipped bbox, for the debug markers */
} quad_setup;

static void setup_quad(const dtro_ctx *c, v2 mn, v2 mx, const float transform[7], quad_setup *q)
{
	/* TransformRectPoints (:378-393) */
	float dimw = mx.x - mn.x, dimh = mx.y - mn.y;
	v2 origin  = {mn.x + (transform[1] * dimw), mn.y + (transform[2] * dimh)};
	v2 in[4]   = {{mn.x - origin.x, mn.y - origin.y},
	              {mx.x - origin.x, mn.y - origin.y},
	              {mx.x - origin.x, mx.y - origin.y},
	              {mn.x - origin.x, mx.y - origin.y}};
	basis2 bs = make_basis(transform[0], transform[4], transform[5]);
	for (int i = 0; i < 4; i++) q->p[i] = apply_basis(origin, bs, in[i]);
	float bminx = q->p[0].x, bminy = q->p[0].y, bmaxx = q->p[0].x, bmaxy = q->p[0].y;
	for (int i = 1; i < 4; i++)
	{
		bminx = ORA_MIN(bminx, q->p[i].x); bminy = ORA_MIN(bminy, q->p[i].y);
		bmaxx = ORA_MAX(bmaxx, q->p[i].x); bmaxy = ORA_MAX(bmaxy, q->p[i].y);
	}
	q->bminx = bminx; q->bminy = bminy; q->bmaxx = bmaxx; q->bmaxy = bmaxy;
	/* clip to (0,0)-(W,H) (:436-440,1628-1632) */
	float clipw = (float)c->width - (float)0, cliph = (float)c->height - (float)0;
	float cmaxx = ORA_MIN(bmaxx, clipw), cmaxy = ORA_MIN(bmaxy, cliph);
	q->cminx = ORA_MAX(0.0f, bminx);
	q->cminy = ORA_MAX(0.0f, bminy);
	q->csizew = cmaxx - q->cminx;
	q->csizeh = cmaxy - q->cminy;
}

/* 4-edge inside test: dot(P - p_i, p_{i+1} - p_i) >= 0 (:456-470,1653-1666) */
static int quad_inside(const quad_setup *q, int x, int y)
{
	for (int i = 0; i < 4; i++)
	{
		v2 o = q->p[i], n = q->p[(i + 1) % 4];
		float ax = n.x - o.x, ay = n.y - o.y;
		float tx = (float)x - o.x, ty = (float)y - o.y;
		float d = 0;
		d += (tx * ax);
		d += (ty * ay);
		if (d < 0) return 0;
	}
	return 1;
}

static uint32_t fetch_texel(const uint8_t *tex, int texW, int x, int y)
{
	uint32_t t;
	memcpy(&t, tex + ((size_t)x * 4 + (size_t)y * (size_t)texW * 4), 4);
	return t;
}

static rgba unpack_texel_linear(uint32_t t)
{
	const float INV_255 = 1.0f / 255.0f; /* `V4 *= DTRRENDER_INV_255` folds to one constant (:1734) */
	rgba o;
	o.a = (float)(t >> 24) * INV_255;
	o.b = (float)((t >> 16) & 0xFF) * INV_255;
	o.g = (float)((t >> 8) & 0xFF) * INV_255;
	o.r = (float)((t >> 0) & 0xFF) * INV_255;
	o.r = o.r * o.r; o.g = o.g * o.g; o.b = o.b * o.b;
	return o;
}

static float lerpf(float a, float t, float b) { return a + (b - a) * t; } /* dqn.h:2301-2317 */

/* ------------------------------------------------------------------------------------------ */
/* public entry points                                                                        */
/* ------------------------------------------------------------------------------------------ */
const char *dtro_kind(void) { return "port"; }

dtro_ctx *dtro_create(int width, int height)
{
	dtro_ctx *c = (dtro_ctx *)calloc(1, sizeof(*c));
	if (!c) return NULL;
	size_t n  = (size_t)width * (size_t)height;
	c->width  = width;
	c->height = height;
	c->color  = (uint32_t *)calloc(n, 4);
	c->z      = (float *)malloc(n * sizeof(float));
	dtro_reset_z(c);
	return c;
}

void dtro_destroy(dtro_ctx *c)
{
	if (!c) return;
	free(c->color);
	free(c->z);
	free(c);
}

uint32_t *dtro_color(dtro_ctx *c) { return c->color; }
float *dtro_zbuffer(dtro_ctx *c) { return c->z; }

void dtro_reset_z(dtro_ctx *c)
{
	size_t n = (size_t)c->width * (size_t)c->height;
	for (size_t i = 0; i < n; i++) c->z[i] = -FLT_MAX; /* DQN_F32_MIN, DTRenderer.cpp:967-978 */
}

uint64_t dtro_counter(dtro_ctx *c, int which) { return c->counters[which ? 1 : 0]; }
void dtro_reset_counters(dtro_ctx *c) { c->counters[0] = c->counters[1] = 0; }

void dtro_clear(dtro_ctx *c, const float rgb[3])
{
	/* :1793-1815 -- truncating conversion, colour only, z untouched */
	float r = rgb[0] * 255.0f, g = rgb[1] * 255.0f, b = rgb[2] * 255.0f;
	uint32_t px = (uint32_t)(((int32_t)0 << 24) | ((int32_t)r << 16) | ((int32_t)g << 8) | ((int32_t)b << 0));
	size_t n = (size_t)c->width * (size_t)c->height;
	for (size_t i = 0; i < n; i++) c->color[i] = px;
}

static const float NO_UV[6] = {0, 0, 0, 0, 0, 0};

void dtro_triangle(dtro_ctx *c, const float p[9], const float color[4], const float transform[7])
{
	draw_triangle(c, p, NO_UV, NULL, 0, 0, DTRO_SHADE_FULLBRIGHT, NULL, NULL, color, transform);
}

void dtro_triangles(dtro_ctx *c, int n, const float *p, const float *color, const float transform[7])
{
	for (int i = 0; i < n; i++) dtro_triangle(c, p + 9 * (size_t)i, color + 4 * (size_t)i, transform);
}

void dtro_textured_triangle(dtro_ctx *c, const float p[9], const float uv[6], const uint8_t *tex,
                            int texW, int texH, const float color[4], const float transform[7])
{
	draw_triangle(c, p, uv, tex, texW, texH, DTRO_SHADE_FULLBRIGHT, NULL, NULL, color, transform);
}

void dtro_mesh(dtro_ctx *c, const float *vertexes, int numVertexes, const float *texUV,
               int numTexUV, const float *normals, int numNormals, const int32_t *faces,
               int numFaces, const uint8_t *tex, int texW, int texH, int lightMode,
               const float lightVector[3], const float lightColor[4], const float pos[3],
               const float transform[7])
{
	(void)numVertexes; (void)numTexUV; (void)numNormals;
	static const float TRI_TRANSFORM[7] = {0, 0.33f, 0.33f, 0.33f, 1, 1, 1}; /* DTRendererRender.h:41-47 */
	mat4 M = mesh_matrix(c->width, c->height, pos, transform);
	for (int i = 0; i < numFaces; i++)
	{
		const int32_t *f = faces + 9 * (size_t)i;
		float p[9], uv[6];
		v3 nrm[3];
		for (int k = 0; k < 3; k++)
		{
			float v[4];
			mat4_mulv4(&M, vertexes + 4 * (size_t)f[k], v);
			float inv = 1.0f / v[3]; /* `xyz / w` is reciprocal-multiply (dqn.h:787) */
			v[0] = v[0] * inv; v[1] = v[1] * inv; v[2] = v[2] * inv;
			p[3 * k + 0] = (float)(int32_t)(v[0] + 0.5f); /* pixel snap (:1485-1490) */
			p[3 * k + 1] = (float)(int32_t)(v[1] + 0.5f);
			p[3 * k + 2] = v[2];
			uv[2 * k + 0] = texUV[3 * (size_t)f[3 + k] + 0];
			uv[2 * k + 1] = texUV[3 * (size_t)f[3 + k] + 1];
			nrm[k].x = normals[3 * (size_t)f[6 + k] + 0];
			nrm[k].y = normals[3 * (size_t)f[6 + k] + 1];
			nrm[k].z = normals[3 * (size_t)f[6 + k] + 2];
		}
		draw_triangle(c, p, uv, tex, texW, texH, lightMode, lightVector, nrm, lightColor, TRI_TRANSFORM);
	}
}

void dtro_rectangle(dtro_ctx *c, const float mn[2], const float mx[2], const float color[4],
                    const float transform[7])
{
	rgba col = to_linear_premul(color);
	quad_setup q;
	v2 a = {mn[0], mn[1]}, b = {mx[0], mx[1]};
	setup_quad(c, a, b, transform, &q);
	if (transform[0] != 0)
	{
		/* the rotated loop swaps w/h bounds (:450-453) -- part of the contract */
		for (int y = 0; (float)y < q.csizew; y++)
		{
			int by = (int)q.cminy + y;
			for (int x = 0; (float)x < q.csizeh; x++)
			{
				int bx = (int)q.cminx + x;
				if (quad_inside(&q, bx, by)) blend_pixel(c, bx, by, col);
			}
		}
	}
	else
	{
		for (int y = 0; (float)y < q.csizeh; y++)
		{
			int by = (int)q.cminy + y;
			for (int x = 0; (float)x < q.csizew; x++)
			{
				int bx = (int)q.cminx + x;
				blend_pixel(c, bx, by, col);
			}
		}
	}
}

void dtro_bitmap(dtro_ctx *c, const uint8_t *tex, int texW, int texH, const float pos[2],
                 const float transform[7], const float color[4])
{
	if (!tex) return;
	v2 mn = {pos[0], pos[1]};
	v2 mx = {pos[0] + (float)texW, pos[1] + (float)texH};
	quad_setup q;
	setup_quad(c, mn, mx, transform, &q);
	rgba col = to_linear_premul(color);

	v2 basis = q.p[0];
	v2 xa = {q.p[1].x - basis.x, q.p[1].y - basis.y};
	v2 ya = {q.p[3].x - basis.x, q.p[3].y - basis.y};
	float tx0 = xa.x - 0, ty0 = xa.y - 0;
	float invx = 1 / ((tx0 * tx0) + (ty0 * ty0));
	tx0 = ya.x - 0; ty0 = ya.y - 0;
	float invy = 1 / ((tx0 * tx0) + (ty0 * ty0));
	int nx = (int)q.csizew, ny = (int)q.csizeh;
	for (int y = 0; y < ny; y++)
	{
		int by = (int)q.cminy + y;
		for (int x = 0; x < nx; x++)
		{
			int bx = (int)q.cminx + x;
			if (!quad_inside(&q, bx, by)) continue;
			float px = (float)bx - basis.x, py = (float)by - basis.y;
			float du = 0; du += (px * xa.x); du += (py * xa.y);
			float dv = 0; dv += (px * ya.x); dv += (py * ya.y);
			float u = clampf(du * invx, 0.0f, 1.0f);
			float v = clampf(dv * invy, 0.0f, 1.0f);
			float txf = u * (float)(texW - 1);
			float tyf = v * (float)(texH - 1);
			int tx = (int)txf, ty = (int)tyf;
			float fx = txf - (float)tx, fy = tyf - (float)ty;
			int tx1 = ORA_MIN(tx + 1, texW - 1), ty1 = ORA_MIN(ty + 1, texH - 1);
			rgba c1 = unpack_texel_linear(fetch_texel(tex, texW, tx, ty));
			rgba c2 = unpack_texel_linear(fetch_texel(tex, texW, tx1, ty));
			rgba c3 = unpack_texel_linear(fetch_texel(tex, texW, tx, ty1));
			rgba c4 = unpack_texel_linear(fetch_texel(tex, texW, tx1, ty1));
			rgba c12 = {lerpf(c1.r, fx, c2.r), lerpf(c1.g, fx, c2.g), lerpf(c1.b, fx, c2.b), lerpf(c1.a, fx, c2.a)};
			rgba c34 = {lerpf(c3.r, fx, c4.r), lerpf(c3.g, fx, c4.g), lerpf(c3.b, fx, c4.b), lerpf(c3.a, fx, c4.a)};
			rgba bl  = {lerpf(c12.r, fy, c34.r), lerpf(c12.g, fy, c34.g), lerpf(c12.b, fy, c34.b), lerpf(c12.a, fy, c34.a)};
			bl.a = bl.a * col.a;
			bl.r = bl.r * col.r;
			bl.g = bl.g * col.g;
			bl.b = bl.b * col.b;
			blend_pixel(c, bx, by, bl);
		}
	}
}

void dtro_line(dtro_ctx *c, const int32_t a_in[2], const int32_t b_in[2], const float color[4])
{
	/* DTRRender_Line (:294-356): integer DDA, x-major after the optional axis swap */
	rgba col = to_linear_premul(color);
	int ax = a_in[0], ay = a_in[1], bx = b_in[0], by = b_in[1];
	int steep = abs(ax - bx) < abs(ay - by);
	if (steep) { int t = ax; ax = ay; ay = t; t = bx; bx = by; by = t; }
	if (bx < ax) { int t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
	int rise = by - ay, run = bx - ax;
	int delta = (by > ay) ? 1 : -1;
	int dist = abs(rise) * 2, acc = 0, ny = ay;
	for (int i = 0; i < run; i++)
	{
		int nx = ax + i;
		if (steep) blend_pixel(c, ny, nx, col);
		else       blend_pixel(c, nx, ny, col);
		acc += dist;
		if (acc > run) { ny += delta; acc -= (run * 2); }
	}
}

void dtro_text(dtro_ctx *c, const uint8_t *atlas, int atlasW, int atlasH, const dtro_packedchar *chars,
               int cpMin, int cpMax, const float pos_in[2], const char *text, const float color[4], int len)
{
	/* DTRRender_Text (:193-273) with stbtt_GetPackedQuad(..., align_to_integer = true)
	 * (external/stb_truetype.h:3739-3763) restated inline.  Quirks kept: the range check at :212-216
	 * can never fire; glyph rows are read from fontHeight down to 1 (:253), not fontHeight-1..0. */
	(void)cpMax;
	if (!text || !atlas || !chars) return;
	if (len == -1) len = (int)strlen(text);
	rgba  col = to_linear_premul(color);
	float posx = pos_in[0], posy = pos_in[1];
	const float ipw = 1.0f / (float)atlasW, iph = 1.0f / (float)atlasH;
	for (int index = 0; index < len; index++)
	{
		const dtro_packedchar *b = chars + ((int)text[index] - cpMin);
		const float qx0 = (float)(int)floor((posx + b->xoff) + 0.5f);
		const float qy0 = (float)(int)floor((posy + b->yoff) + 0.5f);
		const float s0 = (float)b->x0 * ipw, t0 = (float)b->y0 * iph, s1 = (float)b->x1 * ipw, t1 = (float)b->y1 * iph;
		posx += b->xadvance;
		const float fminx = s0 * (float)atlasW, fminy = t1 * (float)atlasH; /* fontRect.min */
		const float fmaxx = s1 * (float)atlasW, fmaxy = t0 * (float)atlasH; /* fontRect.max */
		const uint32_t pitch  = (uint32_t)atlasW;
		const uint32_t offset = (uint32_t)(fminx + (fmaxy * (float)pitch));
		const uint8_t *fontPtr = atlas + offset;
		const float    fho = b->yoff2 + b->yoff;
		const int fontWidth = abs((int)(fminx - fmaxx)), fontHeight = abs((int)(fminy - fmaxy));
		for (int y = 0; y < fontHeight; y++)
			for (int x = 0; x < fontWidth; x++)
			{
				const int     yOffset = fontHeight - y;
				const uint8_t srcA    = fontPtr[x + (yOffset * (int)pitch)];
				if (srcA == 0) continue;
				const float srcANorm = (float)srcA / 255.0f;
				rgba r = {col.r * srcANorm, col.g * srcANorm, col.b * srcANorm, col.a * srcANorm};
				const int actualX = (int)(qx0 + (float)x);
				const int actualY = (int)((qy0 + (float)y) - fho);
				blend_pixel(c, actualX, actualY, r);
			}
	}
}

void dtro_premultiply_bitmap(uint32_t *pixels, int count)
{
	/* DTRAsset_LoadBitmap's per-pixel pass (DTRendererAsset.cpp:823-841) with
	 * DTRRender_PreMultiplyAlphaSRGB1WithLinearConversion (DTRendererRender.cpp:113-121):
	 * byte * (1/255) [reciprocal multiply: the macro is the operand of operator*=], square,
	 * times alpha, sqrtf (0 stays 0), * 255, truncate.  Alpha only goes through *(1/255)*255. */
	const float INV_255 = 1.0f / 255.0f;
	for (int i = 0; i < count; i++)
	{
		const uint32_t pixel = pixels[i];
		float ch[4] = {(float)(pixel & 0xFF), (float)((pixel >> 8) & 0xFF), (float)((pixel >> 16) & 0xFF), (float)(pixel >> 24)};
		for (int k = 0; k < 4; k++) ch[k] = ch[k] * INV_255;
		for (int k = 0; k < 3; k++)
		{
			float v = ch[k] * ch[k];
			v       = v * ch[3];
			ch[k]   = (v == 0) ? 0 : sqrtf(v);
		}
		for (int k = 0; k < 4; k++) ch[k] = ch[k] * 255.0f;
		pixels[i] = ((uint32_t)ch[3] << 24) | ((uint32_t)ch[2] << 16) | ((uint32_t)ch[1] << 8) | (uint32_t)ch[0];
	}
}
