// Stub of tinyexr (saveImage() in materials/Texture.h references it; never called by the oracle).
#pragma once
#define TINYEXR_SUCCESS 0
inline int SaveEXR(const float*, int, int, int, int, const char*, const char**) { return -1; }
inline void FreeEXRErrorMessage(const char*) {}
