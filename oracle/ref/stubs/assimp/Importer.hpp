#pragma once
#include "assimp/scene.h"
namespace Assimp { class Importer {}; }
