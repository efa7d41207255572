// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/dtro.h).
//
// Headless Linux build of the UNMODIFIED reference scalar render path.  The reference
// sources are compiled from where they lie (-I /root/reference/src); nothing is copied.
// Include order follows the reference's unity build (UnityBuild/UnityBuild.cpp:1-4) minus
// the two TUs the path does not need (DTRenderer.cpp, DTRendererAsset.cpp -- the latter does
// not compile under g++ anyway).  The only platform pieces supplied here are the ones
// DTRendererPlatform.h leaves opaque: PlatformLock, PlatformJobQueue, the PlatformAPI
// function table and the PlatformFlags global (DTRenderer.cpp:11).
//
// Variant: -DDTRO_NO_MARKERS turns DTR_DEBUG_RENDER (DTRendererDebug.h:8) off for the render
// TU so Rectangle/Bitmap do not overlay their debug bounding-box/basis lines
// (DTRendererRender.cpp:492-512,1783-1790).  Without it the build is the reference default.
#include "DTRenderer.h"
#include "DTRendererDebug.cpp"
#ifdef DTRO_NO_MARKERS
#undef DTR_DEBUG_RENDER
#define DTR_DEBUG_RENDER 0
#endif
#include "DTRendererRender.cpp"
#define DQN_IMPLEMENTATION
#include "dqn.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "dtro.h"

PlatformFlags globalDTRPlatformFlags;
struct PlatformLock { int unused; };
struct PlatformJobQueue { int unused; };

static void NoopLockAcquire(PlatformLock *const) {}
static void NoopLockRelease(PlatformLock *const) {}

struct dtro_ctx
{
	DTRRenderBuffer  rb;
	DTRRenderContext ctx;
	PlatformAPI      api;
	PlatformLock     lock;
	PlatformJobQueue queue;
	DqnMemStack      tempStack;
};

static DTRRenderTransform MakeTransform(const float t[7])
{
	DTRRenderTransform r = {};
	r.rotation = t[0];
	r.anchor   = DqnV3_3f(t[1], t[2], t[3]);
	r.scale    = DqnV3_3f(t[4], t[5], t[6]);
	return r;
}

static DTRBitmap MakeBitmap(const uint8_t *tex, int w, int h)
{
	DTRBitmap b     = {};
	b.memory        = (u8 *)tex;
	b.dim           = DqnV2i_2i(w, h);
	b.bytesPerPixel = 4;
	return b;
}

extern "C" {

const char *dtro_kind(void) { return "reference"; }

dtro_ctx *dtro_create(int width, int height)
{
	dtro_ctx *c = (dtro_ctx *)calloc(1, sizeof(*c));
	if (!c) return NULL;
	size_t n             = (size_t)width * (size_t)height;
	c->rb.width          = width;
	c->rb.height         = height;
	c->rb.bytesPerPixel  = 4;
	c->rb.renderLock     = &c->lock;
	c->rb.memory         = (volatile u8 *)calloc(n, 4);
	c->rb.zBuffer        = (volatile f32 *)malloc(n * sizeof(f32));
	// The per-pixel lock does a 4-byte CAS on a 1-byte table entry (DTRendererRender.cpp:1020),
	// so pad the tail even though multithread is off here.
	c->rb.pixelLockTable = (volatile bool *)calloc(n + 4, 1);
	c->api.LockAcquire   = NoopLockAcquire;
	c->api.LockRelease   = NoopLockRelease;
	DqnMemStack_Init(&c->tempStack, DQN_MEGABYTE(1), true);

	c->ctx.renderBuffer = &c->rb;
	c->ctx.tempStack    = &c->tempStack;
	c->ctx.api          = &c->api;
	c->ctx.jobQueue     = &c->queue; // DTRRender_Mesh returns early on NULL (:1402)
	c->ctx.multithread  = false;     // the oracle is the in-order single-thread result

	globalDTRPlatformFlags.canUseSSE2  = false; // scalar SlowTriangle is the oracle (:1327-1336)
	globalDTRPlatformFlags.canUseRdtsc = false;
	dtro_reset_z(c);
	return c;
}

void dtro_destroy(dtro_ctx *c)
{
	if (!c) return;
	free((void *)c->rb.memory);
	free((void *)c->rb.zBuffer);
	free((void *)c->rb.pixelLockTable);
	DqnMemStack_Free(&c->tempStack);
	free(c);
}

uint32_t *dtro_color(dtro_ctx *c) { return (uint32_t *)c->rb.memory; }
float *dtro_zbuffer(dtro_ctx *c) { return (float *)c->rb.zBuffer; }

void dtro_reset_z(dtro_ctx *c)
{
	size_t n = (size_t)c->rb.width * (size_t)c->rb.height;
	for (size_t i = 0; i < n; i++) c->rb.zBuffer[i] = DQN_F32_MIN;
}

uint64_t dtro_counter(dtro_ctx *, int which)
{
	return globalDebug.counter[which == DTRO_COUNTER_SETPIXELS ? DTRDebugCounter_SetPixels
	                                                          : DTRDebugCounter_RenderTriangle];
}

void dtro_reset_counters(dtro_ctx *)
{
	for (int i = 0; i < DTRDebugCounter_Count; i++) globalDebug.counter[i] = 0;
}

void dtro_clear(dtro_ctx *c, const float rgb[3])
{
	DTRRender_Clear(c->ctx, DqnV3_3f(rgb[0], rgb[1], rgb[2]));
}

void dtro_triangle(dtro_ctx *c, const float p[9], const float color[4], const float transform[7])
{
	DTRRender_Triangle(c->ctx, DqnV3_3f(p[0], p[1], p[2]), DqnV3_3f(p[3], p[4], p[5]),
	                   DqnV3_3f(p[6], p[7], p[8]),
	                   DqnV4_4f(color[0], color[1], color[2], color[3]), MakeTransform(transform));
}

void dtro_triangles(dtro_ctx *c, int n, const float *p, const float *color, const float transform[7])
{
	for (int i = 0; i < n; i++) dtro_triangle(c, p + 9 * (size_t)i, color + 4 * (size_t)i, transform);
}

void dtro_textured_triangle(dtro_ctx *c, const float p[9], const float uv[6], const uint8_t *tex,
                            int texW, int texH, const float color[4], const float transform[7])
{
	DTRBitmap bmp = MakeBitmap(tex, texW, texH);
	DTRRender_TexturedTriangle(c->ctx, DqnV3_3f(p[0], p[1], p[2]), DqnV3_3f(p[3], p[4], p[5]),
	                           DqnV3_3f(p[6], p[7], p[8]), DqnV2_2f(uv[0], uv[1]),
	                           DqnV2_2f(uv[2], uv[3]), DqnV2_2f(uv[4], uv[5]), tex ? &bmp : NULL,
	                           DqnV4_4f(color[0], color[1], color[2], color[3]),
	                           MakeTransform(transform));
}

void dtro_mesh(dtro_ctx *c, const float *vertexes, int numVertexes, const float *texUV,
               int numTexUV, const float *normals, int numNormals, const int32_t *faces,
               int numFaces, const uint8_t *tex, int texW, int texH, int lightMode,
               const float lightVector[3], const float lightColor[4], const float pos[3],
               const float transform[7])
{
	DTRMesh mesh    = {};
	mesh.vertexes   = (DqnV4 *)vertexes;
	mesh.numVertexes = (u32)numVertexes;
	mesh.texUV      = (DqnV3 *)texUV;
	mesh.numTexUV   = (u32)numTexUV;
	mesh.normals    = (DqnV3 *)normals;
	mesh.numNormals = (u32)numNormals;
	mesh.numFaces   = (u32)numFaces;
	mesh.faces      = (DTRMeshFace *)calloc((size_t)numFaces, sizeof(DTRMeshFace));
	for (int i = 0; i < numFaces; i++)
	{
		i32 *f                     = (i32 *)(faces + 9 * (size_t)i);
		mesh.faces[i].vertexIndex  = f + 0;
		mesh.faces[i].numVertexIndex = 3;
		mesh.faces[i].texIndex     = f + 3;
		mesh.faces[i].numTexIndex  = 3;
		mesh.faces[i].normalIndex  = f + 6;
		mesh.faces[i].numNormalIndex = 3;
	}
	mesh.tex = MakeBitmap(tex, texW, texH);

	DTRRenderLight light = {};
	light.mode   = (enum DTRRenderShadingMode)lightMode;
	light.vector = DqnV3_3f(lightVector[0], lightVector[1], lightVector[2]);
	light.color  = DqnV4_4f(lightColor[0], lightColor[1], lightColor[2], lightColor[3]);

	DTRRender_Mesh(c->ctx, c->ctx.jobQueue, &mesh, light, DqnV3_3f(pos[0], pos[1], pos[2]),
	               MakeTransform(transform));
	free(mesh.faces);
}

void dtro_rectangle(dtro_ctx *c, const float min[2], const float max[2], const float color[4],
                    const float transform[7])
{
	DTRRender_Rectangle(c->ctx, DqnV2_2f(min[0], min[1]), DqnV2_2f(max[0], max[1]),
	                    DqnV4_4f(color[0], color[1], color[2], color[3]), MakeTransform(transform));
}

void dtro_bitmap(dtro_ctx *c, const uint8_t *tex, int texW, int texH, const float pos[2],
                 const float transform[7], const float color[4])
{
	DTRBitmap bmp = MakeBitmap(tex, texW, texH);
	DTRRender_Bitmap(c->ctx, &bmp, DqnV2_2f(pos[0], pos[1]), MakeTransform(transform),
	                 DqnV4_4f(color[0], color[1], color[2], color[3]));
}

void dtro_line(dtro_ctx *c, const int32_t a[2], const int32_t b[2], const float color[4])
{
	DTRRender_Line(c->ctx, DqnV2i_2i(a[0], a[1]), DqnV2i_2i(b[0], b[1]),
	               DqnV4_4f(color[0], color[1], color[2], color[3]));
}

void dtro_text(dtro_ctx *c, const uint8_t *atlas, int atlasW, int atlasH, const dtro_packedchar *chars,
               int cpMin, int cpMax, const float pos[2], const char *text, const float color[4], int len)
{
	static_assert(sizeof(dtro_packedchar) == sizeof(stbtt_packedchar), "dtro_packedchar mirrors stbtt_packedchar");
	DTRFont font        = {};
	font.bitmap         = (u8 *)atlas;
	font.bitmapDim      = DqnV2i_2i(atlasW, atlasH);
	font.codepointRange = DqnV2i_2i(cpMin, cpMax);
	font.sizeInPt       = 0;
	font.atlas          = (stbtt_packedchar *)chars;
	DTRRender_Text(c->ctx, font, DqnV2_2f(pos[0], pos[1]), text, DqnV4_4f(color[0], color[1], color[2], color[3]), len);
}

void dtro_premultiply_bitmap(uint32_t *pixels, int count)
{
	for (int i = 0; i < count; i++)
	{
		// loop body of DTRAsset_LoadBitmap (DTRendererAsset.cpp:823-841), the reference's own helpers
		u32   pixel = pixels[i];
		DqnV4 color = {};
		color.a     = (f32)(pixel >> 24);
		color.b     = (f32)((pixel >> 16) & 0xFF);
		color.g     = (f32)((pixel >> 8) & 0xFF);
		color.r     = (f32)((pixel >> 0) & 0xFF);

		DqnV4 preMulColor = color;
		preMulColor *= DTRRENDER_INV_255;
		preMulColor = DTRRender_PreMultiplyAlphaSRGB1WithLinearConversion(preMulColor);
		preMulColor *= 255.0f;

		pixels[i] = (((u32)preMulColor.a << 24) | ((u32)preMulColor.b << 16) | ((u32)preMulColor.g << 8) |
		             ((u32)preMulColor.r << 0));
	}
}

// The reference's own number parsers (dqn.h:3335-3454), which DTRAsset_LoadWavefrontObj feeds every
// vertex and index through: exported so that the host-side .obj loader's restatement can be pinned.
float   dtro_ref_strtof32(const char *buf, int len) { return Dqn_StrToF32(buf, len); }
int64_t dtro_ref_strtoi64(const char *buf, int len) { return Dqn_StrToI64(buf, len); }

} // extern "C"
