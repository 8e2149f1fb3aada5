// Stand-in: everything the reference uses lives in glm/glm.hpp (see ../../README.md)
#pragma once
#include "../glm.hpp"
