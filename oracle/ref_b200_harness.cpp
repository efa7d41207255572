// ref_b200_harness.cpp -- TEST INFRASTRUCTURE ONLY: the drop-in proof.
//
// Same headless harness and the SAME reference types as ref_harness.cpp, but every draw call goes
// through the host-side mirror dtrenderer_b200/host/DTRRenderB200.h -> C ABI -> CUDA.  It is built
// here (needs the reference HEADERS for its types; no reference .cpp on the draw path) into
// oracle/_ref/libdtr_ref_b200.so, ships to the GPU box as a binary, and tests/test_gpu_dropin.py
// compares its frames with oracle/_ref/libdtr_ref.so (the reference's own CPU draw path).
#include "DTRenderer.h"
#include "DTRendererRender.h"
#define DQN_IMPLEMENTATION
#include "dqn.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "DTRRenderB200.h"
#include "dtro.h"

PlatformFlags globalDTRPlatformFlags;
struct PlatformLock { int unused; };
struct PlatformJobQueue { int unused; };

struct dtro_ctx
{
	DTRRenderBuffer  rb;
	DTRRenderContext ctx;
	PlatformAPI      api;
	PlatformLock     lock;
	PlatformJobQueue queue;
	DqnMemStack      tempStack;
	bool             inFrame;
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

static void Begin(dtro_ctx *c)
{
	if (!c->inFrame)
	{
		DTRRenderB200_BeginFrame(&c->rb);
		c->inFrame = true;
	}
}

static void Sync(dtro_ctx *c)
{
	// hand the frame back to the host buffers, exactly where the app would present it
	if (c->inFrame) DTRRenderB200_EndFrame(&c->rb);
}

extern "C" {

const char *dtro_kind(void) { return "reference-api-on-b200"; }

dtro_ctx *dtro_create(int width, int height)
{
	dtro_ctx *c = (dtro_ctx *)calloc(1, sizeof(*c));
	if (!c) return NULL;
	size_t n            = (size_t)width * (size_t)height;
	c->rb.width         = width;
	c->rb.height        = height;
	c->rb.bytesPerPixel = 4;
	c->rb.renderLock    = &c->lock;
	c->rb.memory        = (volatile u8 *)calloc(n, 4);
	c->rb.zBuffer       = (volatile f32 *)malloc(n * sizeof(f32));
	c->rb.pixelLockTable = (volatile bool *)calloc(n + 4, 1);
	DqnMemStack_Init(&c->tempStack, DQN_MEGABYTE(1), true);
	c->ctx.renderBuffer = &c->rb;
	c->ctx.tempStack    = &c->tempStack;
	c->ctx.api          = &c->api;
	c->ctx.jobQueue     = &c->queue;
	c->ctx.multithread  = false;
	for (size_t i = 0; i < n; i++) c->rb.zBuffer[i] = DQN_F32_MIN;
	if (!DTRB200_Bind(&c->rb))
	{
		free(c);
		return NULL;
	}
	return c;
}

void dtro_destroy(dtro_ctx *c)
{
	if (!c) return;
	DTRB200Binding *b = DTRB200_Bind(&c->rb);
	if (b) dtr_b200_destroy(b->ctx);
	DTRB200_Bindings().erase(&c->rb);
	free((void *)c->rb.memory);
	free((void *)c->rb.zBuffer);
	free((void *)c->rb.pixelLockTable);
	DqnMemStack_Free(&c->tempStack);
	free(c);
}

uint32_t *dtro_color(dtro_ctx *c) { Sync(c); return (uint32_t *)c->rb.memory; }
float *dtro_zbuffer(dtro_ctx *c) { Sync(c); return (float *)c->rb.zBuffer; }

void dtro_reset_z(dtro_ctx *c)
{
	size_t n = (size_t)c->rb.width * (size_t)c->rb.height;
	for (size_t i = 0; i < n; i++) c->rb.zBuffer[i] = DQN_F32_MIN;
	DTRRenderB200_BeginFrame(&c->rb);
	c->inFrame = true;
}

uint64_t dtro_counter(dtro_ctx *c, int which)
{
	dtr_b200_stats s;
	memset(&s, 0, sizeof(s));
	DTRB200Binding *b = DTRB200_Bind(&c->rb);
	if (b)
	{
		dtr_b200_flush(b->ctx);
		dtr_b200_get_stats(b->ctx, &s);
	}
	return which == DTRO_COUNTER_SETPIXELS ? s.setPixels : s.triangles;
}

void dtro_reset_counters(dtro_ctx *c)
{
	DTRB200Binding *b = DTRB200_Bind(&c->rb);
	if (b) dtr_b200_reset_stats(b->ctx);
}

void dtro_clear(dtro_ctx *c, const float rgb[3])
{
	Begin(c);
	DTRRenderB200_Clear(c->ctx, DqnV3_3f(rgb[0], rgb[1], rgb[2]));
}

void dtro_triangle(dtro_ctx *c, const float p[9], const float color[4], const float transform[7])
{
	Begin(c);
	DTRRenderB200_Triangle(c->ctx, DqnV3_3f(p[0], p[1], p[2]), DqnV3_3f(p[3], p[4], p[5]), DqnV3_3f(p[6], p[7], p[8]),
	                       DqnV4_4f(color[0], color[1], color[2], color[3]), MakeTransform(transform));
}

void dtro_triangles(dtro_ctx *c, int n, const float *p, const float *color, const float transform[7])
{
	for (int i = 0; i < n; i++) dtro_triangle(c, p + 9 * (size_t)i, color + 4 * (size_t)i, transform);
}

void dtro_textured_triangle(dtro_ctx *c, const float p[9], const float uv[6], const uint8_t *tex, int texW, int texH,
                            const float color[4], const float transform[7])
{
	Begin(c);
	DTRBitmap bmp = MakeBitmap(tex, texW, texH);
	DTRRenderB200_TexturedTriangle(c->ctx, DqnV3_3f(p[0], p[1], p[2]), DqnV3_3f(p[3], p[4], p[5]),
	                               DqnV3_3f(p[6], p[7], p[8]), DqnV2_2f(uv[0], uv[1]), DqnV2_2f(uv[2], uv[3]),
	                               DqnV2_2f(uv[4], uv[5]), tex ? &bmp : NULL,
	                               DqnV4_4f(color[0], color[1], color[2], color[3]), MakeTransform(transform));
}

void dtro_mesh(dtro_ctx *c, const float *vertexes, int numVertexes, const float *texUV, int numTexUV, const float *normals,
               int numNormals, const int32_t *faces, int numFaces, const uint8_t *tex, int texW, int texH, int lightMode,
               const float lightVector[3], const float lightColor[4], const float pos[3], const float transform[7])
{
	Begin(c);
	// the mesh in ONE block, laid out like DTRAsset_LoadWavefrontObj's model block
	// (DTRendererAsset.cpp:509-578): vertexes | texUV | normals | faces | every face's three index arrays --
	// the mirror then uploads the block as it is and the index table is flattened on the device
	const size_t geometrySize = sizeof(DqnV4) * (size_t)numVertexes, textureSize = sizeof(DqnV3) * (size_t)numTexUV;
	const size_t normalSize = sizeof(DqnV3) * (size_t)numNormals, faceSize = sizeof(DTRMeshFace) * (size_t)numFaces;
	u8 *block = (u8 *)calloc(1, geometrySize + textureSize + normalSize + faceSize + 36 * (size_t)numFaces);
	u8 *at    = block;
	DTRMesh mesh     = {};
	mesh.vertexes    = (DqnV4 *)at;
	at += geometrySize;
	mesh.texUV = (DqnV3 *)at;
	at += textureSize;
	mesh.normals = (DqnV3 *)at;
	at += normalSize;
	mesh.faces = (DTRMeshFace *)at;
	at += faceSize;
	memcpy(mesh.vertexes, vertexes, geometrySize);
	memcpy(mesh.texUV, texUV, textureSize);
	memcpy(mesh.normals, normals, normalSize);
	mesh.numVertexes = (u32)numVertexes;
	mesh.numTexUV    = (u32)numTexUV;
	mesh.numNormals  = (u32)numNormals;
	mesh.numFaces    = (u32)numFaces;
	for (int i = 0; i < numFaces; i++)
	{
		memcpy(at, faces + 9 * (size_t)i, 36);
		i32 *f                       = (i32 *)at;
		at += 36;
		mesh.faces[i].vertexIndex    = f + 0;
		mesh.faces[i].numVertexIndex = 3;
		mesh.faces[i].texIndex       = f + 3;
		mesh.faces[i].numTexIndex    = 3;
		mesh.faces[i].normalIndex    = f + 6;
		mesh.faces[i].numNormalIndex = 3;
	}
	mesh.tex = MakeBitmap(tex, texW, texH);
	DTRRenderLight light = {};
	light.mode   = (enum DTRRenderShadingMode)lightMode;
	light.vector = DqnV3_3f(lightVector[0], lightVector[1], lightVector[2]);
	light.color  = DqnV4_4f(lightColor[0], lightColor[1], lightColor[2], lightColor[3]);
	DTRRenderB200_Mesh(c->ctx, c->ctx.jobQueue, &mesh, light, DqnV3_3f(pos[0], pos[1], pos[2]), MakeTransform(transform));
	// this harness rebuilds the DTRMesh per call, so drop the cache entry keyed by its stack address
	DTRRenderB200_Invalidate(&mesh);
	free(block);
}

void dtro_rectangle(dtro_ctx *c, const float min[2], const float max[2], const float color[4], const float transform[7])
{
	Begin(c);
	DTRRenderB200_Rectangle(c->ctx, DqnV2_2f(min[0], min[1]), DqnV2_2f(max[0], max[1]),
	                        DqnV4_4f(color[0], color[1], color[2], color[3]), MakeTransform(transform));
}

void dtro_bitmap(dtro_ctx *c, const uint8_t *tex, int texW, int texH, const float pos[2], const float transform[7],
                 const float color[4])
{
	Begin(c);
	DTRBitmap bmp = MakeBitmap(tex, texW, texH);
	DTRRenderB200_Bitmap(c->ctx, &bmp, DqnV2_2f(pos[0], pos[1]), MakeTransform(transform),
	                     DqnV4_4f(color[0], color[1], color[2], color[3]));
}

void dtro_line(dtro_ctx *c, const int32_t a[2], const int32_t b[2], const float color[4])
{
	Begin(c);
	DTRRenderB200_Line(c->ctx, DqnV2i_2i(a[0], a[1]), DqnV2i_2i(b[0], b[1]), DqnV4_4f(color[0], color[1], color[2], color[3]));
}

void dtro_text(dtro_ctx *c, const uint8_t *atlas, int atlasW, int atlasH, const dtro_packedchar *chars,
               int cpMin, int cpMax, const float pos[2], const char *text, const float color[4], int len)
{
	Begin(c);
	DTRFont font        = {};
	font.bitmap         = (u8 *)atlas;
	font.bitmapDim      = DqnV2i_2i(atlasW, atlasH);
	font.codepointRange = DqnV2i_2i(cpMin, cpMax);
	font.atlas          = (stbtt_packedchar *)chars;
	DTRRenderB200_Text(c->ctx, font, DqnV2_2f(pos[0], pos[1]), text, DqnV4_4f(color[0], color[1], color[2], color[3]), len);
}

} // extern "C"
