// Stand-in: the post-processing flags the reference passes (rt.hpp:1655-1671). See scene.h.
#pragma once
enum aiPostProcessSteps { aiProcess_JoinIdenticalVertices = 0x2, aiProcess_Triangulate = 0x8, aiProcess_GenNormals = 0x20,
                          aiProcess_GenSmoothNormals = 0x40, aiProcess_PreTransformVertices = 0x100 };
