// DTRAssetB200.h -- a g++-clean Wavefront .obj (subset) loader that produces exactly the in-memory
// mesh DTRAsset_LoadWavefrontObj builds (DTRendererAsset.cpp:190-612), for Linux hosts: the
// reference's own loader does not compile under g++ (DTRendererAsset.cpp:243-244 jumps over
// initialisers).  Header only, depends on dtr_b200.h and libc; no reference header is needed.
//
// Output layout = the reference's "model block" (DTRendererAsset.cpp:509-578), ONE allocation:
//     DqnV4 vertexes[nV] | DqnV3 texUV[nT] | DqnV3 normals[nN] | DTRMeshFace faces[nF] |
//     for every face, in order: i32 vertexIndex[] , i32 texIndex[] , i32 normalIndex[]
// DTRB200ObjMesh's leading members mirror DTRMesh (DTRendererAsset.h:28-42) up to `faces`/`numFaces`,
// and dtr_b200_mesh_face mirrors DTRMeshFace, so the block can be handed to reference code as a
// DTRMesh and to the CUDA module through DTRAssetB200_UploadMesh (dtr_b200_upload_mesh_faces: the
// block is the arena, the per-face arrays are flattened on the device).
//
// Parsing follows the reference statement by statement, quirks included:
//   * `v`, `vt`, `vn` read 2..3 numbers with Dqn_StrToF32's algorithm (dqn.h:3370-3454): an int32 of all
//     digits times 0.1f multiplied up once per decimal; `e+N` exponents are IGNORED, `e-N` shift right;
//     a number list continues over line ends while the next token starts with a digit or '-'
//     (DTRendererAsset.cpp:277-296);
//   * `f` reads v/vt/vn triples, 1-based, stored 0-based; an omitted attribute ("1//3") is skipped, and
//     because the attribute type advances per '/'-or-space separated field, "f 1 2 3" stores ONE vertex
//     whose v/vt/vn are 1,2,3 (DTRendererAsset.cpp:361-414);
//   * `s`, `#` and unknown statements are skipped; `g` skips only itself, so the group name is scanned
//     as the next statement and normally falls to the skip-the-line default (DTRendererAsset.cpp:423-486).
// What the reference asserts on (negative / relative indices, `p` and `l` statements, malformed
// numbers) makes the load fail (false) instead.
#ifndef DTR_ASSET_B200_H
#define DTR_ASSET_B200_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dtr_b200.h"

struct DTRB200ObjMesh
{
	float              *vertexes; // DqnV4[numVertexes]
	uint32_t            numVertexes;
	float              *texUV; // DqnV3[numTexUV]
	uint32_t            numTexUV;
	float              *normals; // DqnV3[numNormals]
	uint32_t            numNormals;
	dtr_b200_mesh_face *faces; // DTRMeshFace[numFaces]
	uint32_t            numFaces;
	// not part of DTRMesh: the model block that holds everything above
	void  *block;
	size_t blockBytes;
};

// Dqn_StrToF32 (dqn.h:3370-3454), restated.  `ok` goes false where the reference hard-asserts.
inline float DTRAssetB200_StrToF32(const char *buf, int bufSize, bool *ok = nullptr)
{
	if (ok) *ok = true;
	if (!buf || bufSize == 0) return 0;
	int  index = 0;
	bool isNegative = false;
	if (buf[index] == '-')
	{
		index++;
		isNegative = true;
	}
	bool     isPastDecimal = false;
	int      numDigitsAfterDecimal = 0;
	uint32_t rawNumber = 0; // the reference's i32 arithmetic, wrapping instead of undefined
	float    digitShiftValue = 1.0f;
	const float digitShiftMultiplier = 0.1f;
	for (int i = index; i < bufSize; i++)
	{
		char ch = buf[i];
		if (ch == '.')
		{
			isPastDecimal = true;
			continue;
		}
		else if (ch == 'e')
		{
			bool positive = true;
			if (i + 1 >= bufSize || !(buf[i + 1] == '-' || buf[i + 1] == '+'))
			{
				if (ok) *ok = false;
				return 0;
			}
			if (buf[i + 1] == '-') positive = false;
			i += 2;
			int  exponentPow = 0;
			bool scientificNotation = false;
			while (i < bufSize)
			{
				scientificNotation = true;
				char e = buf[i];
				if (e >= '0' && e <= '9')
				{
					exponentPow *= 10;
					exponentPow += (e - '0');
				}
				else i = bufSize;
				i++;
			}
			if (!scientificNotation)
			{
				if (ok) *ok = false;
				return 0;
			}
			if (positive) numDigitsAfterDecimal -= exponentPow;
			else numDigitsAfterDecimal += exponentPow;
		}
		else if (ch >= '0' && ch <= '9')
		{
			numDigitsAfterDecimal += (int)isPastDecimal;
			rawNumber *= 10u;
			rawNumber += (uint32_t)(ch - '0');
		}
		else break;
	}
	for (int i = 0; i < numDigitsAfterDecimal; i++) digitShiftValue *= digitShiftMultiplier;
	float result = (float)(int32_t)rawNumber;
	if (numDigitsAfterDecimal > 0) result *= digitShiftValue;
	if (isNegative) result *= -1;
	return result;
}

// Dqn_StrToI64 (dqn.h:3335-3368)
inline int64_t DTRAssetB200_StrToI64(const char *buf, int bufSize)
{
	if (!buf || bufSize == 0) return 0;
	int  index = 0;
	bool isNegative = false;
	if (buf[index] == '-' || buf[index] == '+')
	{
		if (buf[index] == '-') isNegative = true;
		index++;
	}
	else if (!(buf[index] >= '0' && buf[index] <= '9')) return 0;
	uint64_t result = 0;
	for (int i = index; i < bufSize; i++)
	{
		if (buf[i] >= '0' && buf[i] <= '9') result = result * 10u + (uint64_t)(buf[i] - '0');
		else break;
	}
	int64_t r = (int64_t)result;
	return isNegative ? -r : r;
}

inline void DTRAssetB200_FreeMesh(DTRB200ObjMesh *mesh)
{
	if (!mesh) return;
	free(mesh->block);
	memset(mesh, 0, sizeof(*mesh));
}

// The text must be followed by a terminating 0 byte (text[len] == 0): like the reference, the scanner
// looks one character past a statement before it checks the end of the buffer.
inline bool DTRAssetB200_LoadWavefrontObjFromMemory(const char *text, size_t len, DTRB200ObjMesh *mesh)
{
	if (!text || !mesh) return false;
	memset(mesh, 0, sizeof(*mesh));
	struct Face
	{
		std::vector<int32_t> v, t, n;
	};
	std::vector<float> geometry, texture, normal; // 4 / 3 / 3 floats per entry
	std::vector<Face>  faces;
	const char *end = text + len;
	auto isDigit    = [](char c) { return c >= '0' && c <= '9'; };
	auto lower      = [](char c) { return (c >= 'A' && c <= 'Z') ? (char)(c - 'A' + 'a') : c; };
	auto skipBlank  = [&](const char *p) { // FindFirstCharNotLinefeedOrSpace (:176-187)
		while (p < end && (*p == ' ' || *p == '\n' || *p == '\r')) p++;
		return p;
	};
	auto skipLine = [&](const char *p) { // FindFirstNewlineFeedChar (:165-174)
		while (p < end && *p != '\n' && *p != '\r') p++;
		return p;
	};
	for (const char *scan = text; scan < end;)
	{
		switch (lower(*scan))
		{
			case 'v':
			{
				scan++;
				int  type; // 1 geometric, 2 texture, 3 normal
				char id = lower(*scan);
				if (id == ' ') type = 1;
				else if (id == 't' || id == 'n')
				{
					scan++;
					type = (id == 't') ? 2 : 3;
				}
				else return false;
				int   vIndex = 0;
				float v4[4]  = {0, 0, 0, 1.0f};
				for (; scan < end && *scan == ' '; scan++) {}
				for (;;)
				{
					const char *start = scan;
					for (; scan < end && *scan != ' ' && *scan != '\n' && *scan != '\r';)
					{
						if (!(isDigit(*scan) || *scan == '.' || *scan == '-' || *scan == 'e')) return false;
						scan++;
					}
					bool ok = true;
					if (vIndex >= 3) return false; // the reference asserts vIndex < 4 after the increment
					v4[vIndex++] = DTRAssetB200_StrToF32(start, (int)(scan - start), &ok);
					if (!ok) return false;
					scan = skipBlank(scan);
					if (scan >= end) break;
					if (!(isDigit(*scan) || *scan == '-')) break;
				}
				if (vIndex < 2) return false;
				if (type == 1) geometry.insert(geometry.end(), v4, v4 + 4);
				else if (type == 2) texture.insert(texture.end(), v4, v4 + 3);
				else normal.insert(normal.end(), v4, v4 + 3);
			}
			break;
			case 'p':
			case 'l': return false; // the reference asserts (:332-345)
			case 'f':
			{
				scan++;
				scan = skipBlank(scan);
				if (scan >= end) continue;
				Face face;
				int  parsed = 0;
				bool more   = true;
				while (more)
				{
					for (int i = 0; i < 3; i++) // v, vt, vn
					{
						const char *start = scan;
						while (scan < end && isDigit(*scan)) scan++;
						int numLen = (int)(scan - start);
						if (numLen > 0)
						{
							int32_t idx = (int32_t)DTRAssetB200_StrToI64(start, numLen) - 1;
							if (idx < 0) return false; // relative indices are not supported (:386)
							(i == 0 ? face.v : (i == 1 ? face.t : face.n)).push_back(idx);
						}
						if (scan < end) scan++; // the separator, whatever it is
					}
					parsed++;
					scan = skipBlank(scan);
					if (scan >= end || !isDigit(*scan)) more = false;
				}
				if (parsed < 3) return false;
				faces.push_back(face);
			}
			break;
			case 'g':
			{
				// The reference skips the blanks after `g` and then "iterates to the end of the name" with
				// FindFirstCharNotLinefeedOrSpace, which does not move on a name character (:423-452): the
				// group NAME is therefore scanned as the next statement.  Names that start with a letter
				// without a case of its own ("default", "mesh1") fall to the skip-the-line default; a
				// name starting with v, f, g, s, p or l is misread exactly as the reference misreads it.
				scan++;
				scan = skipBlank(scan);
			}
			break;
			case 's':
			{
				scan++;
				scan = skipBlank(scan);
				if (scan < end && isDigit(*scan))
				{
					while (scan < end && *scan != ' ' && *scan != '\n' && *scan != '\r')
					{
						if (!isDigit(*scan)) return false;
						scan++;
					}
				}
				scan = skipBlank(scan);
			}
			break;
			default: // comments and everything unrecognised: to the end of the line (:471-486)
				scan = skipLine(scan);
				scan = skipBlank(scan);
				break;
		}
	}

	// ---- the compact model block (:488-578) -------------------------------------------------------
	const size_t nV = geometry.size() / 4, nT = texture.size() / 3, nN = normal.size() / 3, nF = faces.size();
	size_t       total = nV * 16 + nT * 12 + nN * 12 + nF * sizeof(dtr_b200_mesh_face);
	for (const Face &f : faces) total += (f.v.size() + f.t.size() + f.n.size()) * sizeof(int32_t);
	uint8_t *block = (uint8_t *)calloc(1, total ? total : 1);
	if (!block) return false;
	uint8_t *p     = block;
	mesh->block    = block;
	mesh->blockBytes = total;
	mesh->vertexes = (float *)p;
	p += nV * 16;
	mesh->texUV = (float *)p;
	p += nT * 12;
	mesh->normals = (float *)p;
	p += nN * 12;
	mesh->faces = (dtr_b200_mesh_face *)p;
	p += nF * sizeof(dtr_b200_mesh_face);
	mesh->numVertexes = (uint32_t)nV;
	mesh->numTexUV    = (uint32_t)nT;
	mesh->numNormals  = (uint32_t)nN;
	mesh->numFaces    = (uint32_t)nF;
	if (nV) memcpy(mesh->vertexes, geometry.data(), nV * 16);
	if (nT) memcpy(mesh->texUV, texture.data(), nT * 12);
	if (nN) memcpy(mesh->normals, normal.data(), nN * 12);
	for (size_t i = 0; i < nF; i++)
	{
		const Face         &f  = faces[i];
		dtr_b200_mesh_face &mf = mesh->faces[i];
		auto put = [&](const std::vector<int32_t> &a, const int32_t *&dst, uint32_t &count) {
			dst   = (const int32_t *)p;
			count = (uint32_t)a.size();
			if (!a.empty()) memcpy(p, a.data(), a.size() * sizeof(int32_t));
			p += a.size() * sizeof(int32_t);
		};
		put(f.v, mf.vertexIndex, mf.numVertexIndex);
		put(f.t, mf.texIndex, mf.numTexIndex);
		put(f.n, mf.normalIndex, mf.numNormalIndex);
	}
	return true;
}

inline bool DTRAssetB200_LoadWavefrontObj(const char *path, DTRB200ObjMesh *mesh)
{
	if (!path || !mesh) return false;
	FILE *f = fopen(path, "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long size = ftell(f);
	fseek(f, 0, SEEK_SET);
	if (size < 0)
	{
		fclose(f);
		return false;
	}
	std::vector<char> text((size_t)size + 1, 0);
	size_t got = fread(text.data(), 1, (size_t)size, f);
	fclose(f);
	if (got != (size_t)size) return false;
	return DTRAssetB200_LoadWavefrontObjFromMemory(text.data(), (size_t)size, mesh);
}

// Hand the loaded mesh to the CUDA module: the model block is the arena, the index table is flattened
// on the device (dtr_b200_upload_mesh_faces).
inline int DTRAssetB200_UploadMesh(dtr_b200_ctx *ctx, const DTRB200ObjMesh *mesh, int texId, int *meshId)
{
	if (!mesh || !mesh->block) return DTR_B200_ERR_ARG;
	dtr_b200_mesh_faces_desc d = {mesh->vertexes, mesh->numVertexes, mesh->texUV,  mesh->numTexUV, mesh->normals,
	                              mesh->numNormals, mesh->faces,      mesh->numFaces, mesh->block, mesh->blockBytes};
	return dtr_b200_upload_mesh_faces(ctx, &d, texId, meshId);
}

#endif // DTR_ASSET_B200_H
