#pragma once
enum aiPostProcessSteps { aiProcess_Triangulate = 0x8, aiProcess_FlipUVs = 0x800000, aiProcess_CalcTangentSpace = 0x1, aiProcess_GenNormals = 0x20 };
