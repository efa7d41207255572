/* dtro.h -- C interface shared by BOTH checkers under oracle/:
 *   oracle/_ref/libdtr_ref.so   the UNMODIFIED reference scalar path, compiled from
 *                               /root/reference/src by oracle/Makefile (ref_harness.cpp)
 *   oracle/libdtr_oracle.so     the plain-C restatement (dtr_oracle.c)
 *
 * TEST INFRASTRUCTURE ONLY: nothing in the product path (dtrenderer_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Every entry mirrors one reference draw call (DTRendererRender.h:91-98) with the
 * arguments flattened to plain float arrays:
 *   transform = {rotation, anchor.x, anchor.y, anchor.z, scale.x, scale.y, scale.z}
 *               (DTRRenderTransform, DTRendererRender.h:28-33)
 *   mesh      = vertexes f32[nV*4], texUV f32[nT*3], normals f32[nN*3],
 *               faces i32[nF*9] = {v0,v1,v2, t0,t1,t2, n0,n1,n2}  (DTRendererAsset.h:16-41)
 *   texture   = u8[h*w*4], texel u32 = A<<24|B<<16|G<<8|R, premultiplied
 *               (DTRendererAsset.cpp:825-841); pass NULL for "no texture".
 */
#ifndef DTRO_H
#define DTRO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dtro_ctx dtro_ctx;

enum { DTRO_SHADE_FULLBRIGHT = 0, DTRO_SHADE_FLAT = 1, DTRO_SHADE_GOURAUD = 2 };
enum { DTRO_COUNTER_SETPIXELS = 0, DTRO_COUNTER_TRIANGLES = 1 };

const char *dtro_kind(void); /* "reference" or "port" */

dtro_ctx *dtro_create(int width, int height);
void      dtro_destroy(dtro_ctx *c);
uint32_t *dtro_color(dtro_ctx *c);            /* u32[W*H], 0x00RRGGBB, row 0 = bottom */
float    *dtro_zbuffer(dtro_ctx *c);          /* f32[W*H], larger = nearer */
void      dtro_reset_z(dtro_ctx *c);          /* z = -FLT_MAX (DTRenderer.cpp:967-978) */
uint64_t  dtro_counter(dtro_ctx *c, int which);
void      dtro_reset_counters(dtro_ctx *c);

void dtro_clear(dtro_ctx *c, const float rgb[3]);
void dtro_triangle(dtro_ctx *c, const float p[9], const float color[4], const float transform[7]);
/* n triangles with one shared transform, submitted in order */
void dtro_triangles(dtro_ctx *c, int n, const float *p /*n*9*/, const float *color /*n*4*/,
                    const float transform[7]);
void dtro_textured_triangle(dtro_ctx *c, const float p[9], const float uv[6],
                            const uint8_t *tex, int texW, int texH,
                            const float color[4], const float transform[7]);
void dtro_mesh(dtro_ctx *c,
               const float *vertexes, int numVertexes,
               const float *texUV, int numTexUV,
               const float *normals, int numNormals,
               const int32_t *faces, int numFaces,
               const uint8_t *tex, int texW, int texH,
               int lightMode, const float lightVector[3], const float lightColor[4],
               const float pos[3], const float transform[7]);
void dtro_rectangle(dtro_ctx *c, const float min[2], const float max[2], const float color[4],
                    const float transform[7]);
void dtro_bitmap(dtro_ctx *c, const uint8_t *tex, int texW, int texH, const float pos[2],
                 const float transform[7], const float color[4]);
void dtro_line(dtro_ctx *c, const int32_t a[2], const int32_t b[2], const float color[4]);

/* DTRRender_Text (DTRendererRender.cpp:193-273).  The font is passed flattened: the 1-byte-per-pixel
 * atlas (DTRFont::bitmap, bitmapDim) and one dtro_packedchar per codepoint of [cpMin, cpMax) -- the
 * layout of stbtt_packedchar (external/stb_truetype.h:522-527), which is what DTRFont::atlas holds. */
typedef struct dtro_packedchar
{
	uint16_t x0, y0, x1, y1;
	float    xoff, yoff, xadvance, xoff2, yoff2;
} dtro_packedchar;
void dtro_text(dtro_ctx *c, const uint8_t *atlas, int atlasW, int atlasH, const dtro_packedchar *chars,
               int cpMin, int cpMax, const float pos[2], const char *text, const float color[4], int len);

/* The per-pixel pass of DTRAsset_LoadBitmap (DTRendererAsset.cpp:816-843): straight-alpha RGBA8
 * texels (R in the low byte, as stb_image returns them) -> premultiplied in sRGB space, in place.
 * The reference build calls the renderer's own DTRRender_PreMultiplyAlphaSRGB1WithLinearConversion
 * (DTRendererRender.cpp:113-121) inside the loop restated from the asset file (that file does not
 * compile under g++). */
void dtro_premultiply_bitmap(uint32_t *pixels, int count);

#ifdef __cplusplus
}
#endif
#endif
