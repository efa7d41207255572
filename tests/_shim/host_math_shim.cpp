// Test-only shim: exposes the product's per-draw-call host arithmetic (dtr_host_math.h) to the CPU
// test-suite, so that it can be checked against the reference without a GPU.
#include "dtr_host_math.h"

extern "C" {

uint32_t hm_pack_clear(const float *rgb) { return dtr::pack_clear(rgb); }

// returns 1 when something is drawn; bbox = {x0, y0, x1, y1} half-open pixel bounds, type = PrimType
int hm_rect(int W, int H, const float *mn, const float *mx, float rotation, const float *anchor, const float *scale,
            const float *color, int *bbox, int *type, float *points8)
{
	dtr::PrimRecord rec;
	float           pts[4][2];
	bool            ok = dtr::setup_quad(W, H, mn, mx, rotation, anchor, scale, color, false, 0, 0, 0, &rec, pts, nullptr);
	for (int i = 0; i < 4; i++)
	{
		points8[2 * i]     = pts[i][0];
		points8[2 * i + 1] = pts[i][1];
	}
	if (!ok) return 0;
	bbox[0] = (int)(rec.w[dtr::QW_MIN] & 0xFFFF);
	bbox[1] = (int)(rec.w[dtr::QW_MIN] >> 16);
	bbox[2] = (int)(rec.w[dtr::QW_MAX] & 0xFFFF);
	bbox[3] = (int)(rec.w[dtr::QW_MAX] >> 16);
	*type   = (int)(rec.w[dtr::QW_FLAGS] & dtr::PF_TYPE_MASK);
	return 1;
}

void hm_mesh_matrix(int W, int H, const float *pos, float rotationDegrees, const float *axis, const float *scale, float *out16)
{
	dtr::Mat4 m = dtr::mesh_matrix(W, H, pos, rotationDegrees, axis, scale);
	std::memcpy(out16, m.e, sizeof(m.e));
}
}
