// DTRRenderB200.h -- host-side mirror of DTRenderer's draw-call API on top of the C ABI.
//
// Include this AFTER the reference's own headers (dqn.h, DTRendererRender.h, DTRendererAsset.h):
// it uses the reference's types as they are (DTRRenderContext, DTRRenderTransform, DTRRenderLight,
// DTRMesh, DTRBitmap, DqnV2/V3/V4) and provides one function per draw call with the SAME
// signature, argument meaning and silent-return behaviour (DTRendererRender.h:91-98):
//
//     DTRRender_Clear            -> DTRRenderB200_Clear
//     DTRRender_Triangle         -> DTRRenderB200_Triangle
//     DTRRender_TexturedTriangle -> DTRRenderB200_TexturedTriangle
//     DTRRender_Mesh             -> DTRRenderB200_Mesh
//     DTRRender_Rectangle        -> DTRRenderB200_Rectangle
//     DTRRender_Bitmap           -> DTRRenderB200_Bitmap
//     DTRRender_Line             -> DTRRenderB200_Line
//     DTRRender_Text             -> DTRRenderB200_Text
//
// plus the two frame hooks the app (DTR_Update, DTRenderer.cpp:967-989 / :1097-1102) calls:
//     DTRRenderB200_BeginFrame(renderBuffer)  after the per-frame z-buffer reset
//     DTRRenderB200_EndFrame(renderBuffer)    before the platform presents renderBuffer->memory
//
// The binding owns one dtr_b200_ctx per DTRRenderBuffer (created on first use, on the device chosen
// with DTRB200_SetDevice, default 0) and caches the uploads of DTRMesh / DTRBitmap / DTRFont by
// address AND a cheap content fingerprint (dimensions / counts plus a strided sample of the bytes): an
// asset whose memory is reused for something else is uploaded again.  An asset edited in place in a
// way the sample misses is refreshed with DTRRenderB200_Invalidate(ptr).  No reference source is
// modified or copied; see INTEGRATION.md for the three-line patch that switches the renderer over.
#ifndef DTR_RENDER_B200_H
#define DTR_RENDER_B200_H

#include <map>
#include <vector>

#include "dtr_b200.h"

#ifndef DTR_B200_DEBUG_MARKERS
#define DTR_B200_DEBUG_MARKERS 0
#endif

struct DTRB200Cached
{
	int      id   = -1;
	uint64_t mark = 0; // content fingerprint at upload time
};

struct DTRB200Binding
{
	dtr_b200_ctx                         *ctx = nullptr;
	std::map<const void *, DTRB200Cached> textures; // DTRBitmap::memory -> texId
	std::map<const void *, DTRB200Cached> meshes;   // DTRMesh * -> meshId
	std::map<const void *, DTRB200Cached> fonts;    // DTRFont::bitmap -> fontId
};

inline int &DTRB200_Device()
{
	static int device = 0;
	return device;
}
// Which CUDA device new bindings are created on (one host thread per device).
inline void DTRB200_SetDevice(int device) { DTRB200_Device() = device; }

// FNV-1a over the sizes and at most 4096 bytes sampled evenly from the asset: cheap per draw call.
inline uint64_t DTRB200_Fingerprint(const void *data, size_t bytes, uint64_t seed)
{
	uint64_t       h = 1469598103934665603ull ^ seed;
	const uint8_t *p = (const uint8_t *)data;
	const size_t   step = bytes > 4096 ? bytes / 4096 : 1;
	for (size_t i = 0; i < bytes; i += step) h = (h ^ p[i]) * 1099511628211ull;
	return (h ^ bytes) * 1099511628211ull;
}

inline std::map<const DTRRenderBuffer *, DTRB200Binding> &DTRB200_Bindings()
{
	static std::map<const DTRRenderBuffer *, DTRB200Binding> bindings;
	return bindings;
}

inline DTRB200Binding *DTRB200_Bind(const DTRRenderBuffer *rb)
{
	if (!rb) return nullptr;
	DTRB200Binding &b = DTRB200_Bindings()[rb];
	if (!b.ctx)
	{
		if (dtr_b200_create(DTRB200_Device(), rb->width, rb->height, 1, &b.ctx) != DTR_B200_OK) return nullptr;
		// DTR_B200_DEBUG_MARKERS 1 reproduces the overlay of the reference's default
		// (DTR_DEBUG_RENDER 1) build: bounding boxes, rotated outlines, bitmap corner markers
		dtr_b200_set_debug_markers(b.ctx, DTR_B200_DEBUG_MARKERS);
	}
	return &b;
}

inline dtr_b200_transform DTRB200_Transform(const DTRRenderTransform &t)
{
	dtr_b200_transform r = {t.rotation, {t.anchor.x, t.anchor.y, t.anchor.z}, {t.scale.x, t.scale.y, t.scale.z}};
	return r;
}

inline int DTRB200_Texture(DTRB200Binding *b, const DTRBitmap *bmp)
{
	if (!bmp || !bmp->memory) return -1;
	const size_t   bytes = (size_t)bmp->dim.w * bmp->dim.h * bmp->bytesPerPixel;
	const uint64_t mark  = DTRB200_Fingerprint(bmp->memory, bytes, ((uint64_t)bmp->dim.w << 32) | (uint32_t)bmp->dim.h);
	auto           it    = b->textures.find(bmp->memory);
	if (it != b->textures.end() && it->second.mark == mark) return it->second.id;
	int id = -1;
	if (dtr_b200_upload_texture(b->ctx, bmp->memory, bmp->dim.w, bmp->dim.h, bmp->bytesPerPixel, &id) != DTR_B200_OK) return -1;
	b->textures[bmp->memory] = DTRB200Cached{id, mark};
	return id;
}

// Forget what was uploaded from this address (DTRBitmap::memory, a DTRMesh *, DTRFont::bitmap) in every
// binding: the next draw call that uses it uploads it again.
inline void DTRRenderB200_Invalidate(const void *asset)
{
	for (auto &kv : DTRB200_Bindings())
	{
		kv.second.textures.erase(asset);
		kv.second.meshes.erase(asset);
		kv.second.fonts.erase(asset);
	}
}

// Frame hooks ------------------------------------------------------------------------------------
// hostZ == zBuffer freshly reset by the app: only its reset is mirrored (no upload of 8 MB of -FLT_MAX).
inline void DTRRenderB200_BeginFrame(DTRRenderBuffer *rb)
{
	DTRB200Binding *b = DTRB200_Bind(rb);
	if (b) dtr_b200_begin_frame(b->ctx, 0, nullptr, nullptr);
}

inline void DTRRenderB200_EndFrame(DTRRenderBuffer *rb)
{
	DTRB200Binding *b = DTRB200_Bind(rb);
	if (b) dtr_b200_end_frame(b->ctx, 0, (uint32_t *)rb->memory, (float *)rb->zBuffer);
}

// Draw calls -------------------------------------------------------------------------------------
inline void DTRRenderB200_Clear(DTRRenderContext context, DqnV3 color)
{
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	dtr_b200_clear(b->ctx, color.e);
}

inline void DTRRenderB200_Triangle(DTRRenderContext context, DqnV3 p1, DqnV3 p2, DqnV3 p3, DqnV4 color,
                                   const DTRRenderTransform transform = DTRRender_DefaultTriangleTransform())
{
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	dtr_b200_transform t = DTRB200_Transform(transform);
	dtr_b200_triangle(b->ctx, p1.e, p2.e, p3.e, color.e, &t);
}

inline void DTRRenderB200_TexturedTriangle(DTRRenderContext context, DqnV3 p1, DqnV3 p2, DqnV3 p3, DqnV2 uv1, DqnV2 uv2,
                                           DqnV2 uv3, DTRBitmap *const texture, DqnV4 color,
                                           const DTRRenderTransform transform = DTRRender_DefaultTriangleTransform())
{
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	dtr_b200_transform t = DTRB200_Transform(transform);
	dtr_b200_textured_triangle(b->ctx, p1.e, p2.e, p3.e, uv1.e, uv2.e, uv3.e, DTRB200_Texture(b, texture), color.e, &t);
}

inline void DTRRenderB200_Mesh(DTRRenderContext context, PlatformJobQueue *const jobQueue, DTRMesh *const mesh,
                               DTRRenderLight lighting, const DqnV3 pos, const DTRRenderTransform transform)
{
	// same guard as the reference (DTRendererRender.cpp:1402); the job queue itself is not used
	if (!mesh || !context.renderBuffer || !context.tempStack || !context.api || !jobQueue) return;
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	int            meshId = -1;
	const uint64_t mark   = DTRB200_Fingerprint(mesh->vertexes, sizeof(DqnV4) * (size_t)mesh->numVertexes,
	                                            ((uint64_t)mesh->numFaces << 32) | mesh->numVertexes) ^
	                      DTRB200_Fingerprint(mesh->faces, sizeof(DTRMeshFace) * (size_t)mesh->numFaces, mesh->numNormals);
	auto it = b->meshes.find(mesh);
	if (it != b->meshes.end() && it->second.mark == mark) meshId = it->second.id;
	else
	{
		const int texId = DTRB200_Texture(b, &mesh->tex);
		// DTRAsset_LoadWavefrontObj leaves the whole mesh in ONE memory block -- vertexes, texUV, normals,
		// faces, then every face's three index arrays in order (DTRendererAsset.cpp:509-578).  If this mesh
		// looks like that, the block is uploaded as it is and the index table is flattened on the device
		// (which also verifies that every pointer stays inside the block); otherwise, or if that check
		// fails, the per-face arrays are gathered here.
		static_assert(sizeof(dtr_b200_mesh_face) == sizeof(DTRMeshFace), "dtr_b200_mesh_face mirrors DTRMeshFace");
		bool uploaded = false;
		if (mesh->numFaces)
		{
			const DTRMeshFace &last  = mesh->faces[mesh->numFaces - 1];
			const uint8_t     *begin = (const uint8_t *)mesh->vertexes;
			const uint8_t     *end   = (const uint8_t *)(last.normalIndex + last.numNormalIndex);
			if (last.normalIndex && end > begin && (const uint8_t *)mesh->faces > begin && (const uint8_t *)mesh->faces < end)
			{
				dtr_b200_mesh_faces_desc d = {(const float *)mesh->vertexes, mesh->numVertexes, (const float *)mesh->texUV,
				                              mesh->numTexUV,                (const float *)mesh->normals, mesh->numNormals,
				                              (const dtr_b200_mesh_face *)mesh->faces, mesh->numFaces, begin, (size_t)(end - begin)};
				uploaded = dtr_b200_upload_mesh_faces(b->ctx, &d, texId, &meshId) == DTR_B200_OK;
			}
		}
		if (!uploaded)
		{
			std::vector<int32_t> faces((size_t)mesh->numFaces * 9);
			for (u32 i = 0; i < mesh->numFaces; i++)
			{
				const DTRMeshFace &f = mesh->faces[i];
				if (f.numVertexIndex != 3 || f.numNormalIndex != 3 || f.numTexIndex < 3) return; // reference asserts
				for (int k = 0; k < 3; k++)
				{
					faces[9 * i + k]     = f.vertexIndex[k];
					faces[9 * i + 3 + k] = f.texIndex[k];
					faces[9 * i + 6 + k] = f.normalIndex[k];
				}
			}
			dtr_b200_mesh_desc d = {(const float *)mesh->vertexes, mesh->numVertexes, (const float *)mesh->texUV, mesh->numTexUV,
			                        (const float *)mesh->normals,  mesh->numNormals,  faces.data(),               mesh->numFaces};
			if (dtr_b200_upload_mesh(b->ctx, &d, texId, &meshId) != DTR_B200_OK) return;
		}
		b->meshes[mesh] = DTRB200Cached{meshId, mark};
	}
	dtr_b200_light     l = {(int32_t)lighting.mode, {lighting.vector.x, lighting.vector.y, lighting.vector.z},
	                        {lighting.color.r, lighting.color.g, lighting.color.b, lighting.color.a}};
	dtr_b200_transform t = DTRB200_Transform(transform);
	dtr_b200_mesh(b->ctx, meshId, &l, pos.e, &t);
}

inline void DTRRenderB200_Rectangle(DTRRenderContext context, DqnV2 min, DqnV2 max, DqnV4 color,
                                    const DTRRenderTransform transform = DTRRender_DefaultTransform())
{
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	dtr_b200_transform t = DTRB200_Transform(transform);
	dtr_b200_rectangle(b->ctx, min.e, max.e, color.e, &t);
}

inline void DTRRenderB200_Bitmap(DTRRenderContext context, DTRBitmap *const bitmap, DqnV2 pos,
                                 const DTRRenderTransform transform = DTRRender_DefaultTransform(),
                                 DqnV4 color = DqnV4_4f(1, 1, 1, 1))
{
	if (!bitmap || !bitmap->memory) return; // DTRendererRender.cpp:1601
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	dtr_b200_transform t = DTRB200_Transform(transform);
	dtr_b200_bitmap(b->ctx, DTRB200_Texture(b, bitmap), pos.e, &t, color.e);
}

// DTRFont::atlas is an array of stbtt_packedchar, which dtr_b200_packedchar mirrors field for field.
inline void DTRRenderB200_Text(DTRRenderContext context, const DTRFont font, DqnV2 pos, const char *const text,
                               DqnV4 color = DqnV4_4f(1, 1, 1, 1), i32 len = -1)
{
	if (!text || !font.bitmap || !font.atlas) return; // DTRendererRender.cpp:197-200
	static_assert(sizeof(dtr_b200_packedchar) == sizeof(stbtt_packedchar), "packed char layout");
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	int            id   = -1;
	const uint64_t mark = DTRB200_Fingerprint(font.bitmap, (size_t)font.bitmapDim.w * font.bitmapDim.h,
	                                          ((uint64_t)font.bitmapDim.w << 32) | (uint32_t)font.bitmapDim.h);
	auto           it   = b->fonts.find(font.bitmap);
	if (it != b->fonts.end() && it->second.mark == mark) id = it->second.id;
	else
	{
		if (dtr_b200_upload_font(b->ctx, font.bitmap, font.bitmapDim.w, font.bitmapDim.h, (const dtr_b200_packedchar *)font.atlas,
		                         font.codepointRange.min, font.codepointRange.max, &id) != DTR_B200_OK)
			return;
		b->fonts[font.bitmap] = DTRB200Cached{id, mark};
	}
	dtr_b200_text(b->ctx, id, pos.e, text, color.e, len);
}

inline void DTRRenderB200_Line(DTRRenderContext context, DqnV2i a, DqnV2i b2, DqnV4 color)
{
	DTRB200Binding *b = DTRB200_Bind(context.renderBuffer);
	if (!b) return;
	dtr_b200_line(b->ctx, a.e, b2.e, color.e);
}

#endif // DTR_RENDER_B200_H
