// Test-only shim: exposes the host-side .obj loader (dtrenderer_b200/host/DTRAssetB200.h) to the CPU
// test-suite.  dtr_b200.h is only needed for its plain structs; nothing of the CUDA module is called.
#include "DTRAssetB200.h"

extern "C" {

float   objshim_strtof32(const char *buf, int len, int *ok)
{
	bool  o = true;
	float v = DTRAssetB200_StrToF32(buf, len, &o);
	*ok     = o ? 1 : 0;
	return v;
}
int64_t objshim_strtoi64(const char *buf, int len) { return DTRAssetB200_StrToI64(buf, len); }

// Loads `text` (len bytes + a terminating 0) and returns the mesh through flat arrays the caller sized
// from a first call with NULL outputs: counts = {nV, nT, nN, nF, blockBytes, layoutOk}.
int objshim_load(const char *text, long len, long counts[6], float *vertexes, float *texUV, float *normals, int32_t *faces9,
                 uint32_t *faceCounts3)
{
	DTRB200ObjMesh m;
	if (!DTRAssetB200_LoadWavefrontObjFromMemory(text, (size_t)len, &m)) return 0;
	counts[0] = m.numVertexes;
	counts[1] = m.numTexUV;
	counts[2] = m.numNormals;
	counts[3] = m.numFaces;
	counts[4] = (long)m.blockBytes;
	// the model block's layout (DTRendererAsset.cpp:509-578): arrays back to back, then the per-face index arrays in order
	const uint8_t *b  = (const uint8_t *)m.block;
	bool           ok = (const uint8_t *)m.vertexes == b && (const uint8_t *)m.texUV == b + 16 * (size_t)m.numVertexes &&
	          (const uint8_t *)m.normals == (const uint8_t *)m.texUV + 12 * (size_t)m.numTexUV &&
	          (const uint8_t *)m.faces == (const uint8_t *)m.normals + 12 * (size_t)m.numNormals;
	const uint8_t *p = (const uint8_t *)(m.faces + m.numFaces);
	for (uint32_t i = 0; i < m.numFaces && ok; i++)
	{
		const dtr_b200_mesh_face &f = m.faces[i];
		ok = ok && (const uint8_t *)f.vertexIndex == p;
		p += 4 * (size_t)f.numVertexIndex;
		ok = ok && (const uint8_t *)f.texIndex == p;
		p += 4 * (size_t)f.numTexIndex;
		ok = ok && (const uint8_t *)f.normalIndex == p;
		p += 4 * (size_t)f.numNormalIndex;
	}
	ok        = ok && p == b + m.blockBytes;
	counts[5] = ok ? 1 : 0;
	if (vertexes)
	{
		memcpy(vertexes, m.vertexes, 16 * (size_t)m.numVertexes);
		memcpy(texUV, m.texUV, 12 * (size_t)m.numTexUV);
		memcpy(normals, m.normals, 12 * (size_t)m.numNormals);
		for (uint32_t i = 0; i < m.numFaces; i++)
		{
			const dtr_b200_mesh_face &f = m.faces[i];
			faceCounts3[3 * i]          = f.numVertexIndex;
			faceCounts3[3 * i + 1]      = f.numTexIndex;
			faceCounts3[3 * i + 2]      = f.numNormalIndex;
			for (int k = 0; k < 3; k++)
			{
				faces9[9 * i + k]     = k < (int)f.numVertexIndex ? f.vertexIndex[k] : -1;
				faces9[9 * i + 3 + k] = k < (int)f.numTexIndex ? f.texIndex[k] : -1;
				faces9[9 * i + 6 + k] = k < (int)f.numNormalIndex ? f.normalIndex[k] : -1;
			}
		}
	}
	DTRAssetB200_FreeMesh(&m);
	return 1;
}
}
