// Host API — scene description files; grammar in SceneFile.cpp.
#ifndef HAI719_HOST_SCENEFILE_H
#define HAI719_HOST_SCENEFILE_H
#include <string>
#include "Scene.h"

namespace hai719 {
// Replaces the contents of `scene` with what `filename` describes (paths inside the file are relative to
// scene.asset_root). On error returns false with "<file>:<line>: <what>" in *error; the scene is then unspecified
// but valid. Never calls exit().
bool load_scene_file(Scene &scene, const std::string &filename, std::string *error);
}  // namespace hai719
#endif
