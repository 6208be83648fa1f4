#include "feature_cache.hpp"
#include <cstdio>
#include <cstring>

namespace cvghost {

namespace {
bool read_exact(FILE* f, void* dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, f) == bytes; }

bool read_u8_as_float(FILE* f, std::vector<float>& out, size_t count)
{
    std::vector<uint8_t> tmp(count);
    if (!read_exact(f, tmp.data(), count)) return false;
    out.resize(count);
    for (size_t i = 0; i < count; i++) out[i] = (float)tmp[i];
    return true;
}

bool read_names(FILE* f, std::vector<std::string>& out, int n)
{
    out.clear();
    for (int i = 0; i < n; i++) {
        char buf[65] = { 0 };
        if (!read_exact(f, buf, 64)) return false;
        out.emplace_back(buf);
    }
    return true;
}
}  // namespace

bool FeatureCache::load(const std::string& path, std::string* err)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    bool ok = false;
    do {
        char magic[4]; int32_t hdr[5];
        if (!read_exact(f, magic, 4) || memcmp(magic, "CVGF", 4) != 0) break;
        if (!read_exact(f, hdr, sizeof hdr) || hdr[0] != 1) break;
        n_views = hdr[1]; n_scenes = hdr[2]; n_scales = hdr[3]; n_model_rows = hdr[4];
        if (!read_exact(f, &n_scene_rows, 8)) break;
        view_offsets.resize(n_views + 1); view_model.resize(n_views);
        if (!read_exact(f, view_offsets.data(), view_offsets.size() * 4)) break;
        if (!read_exact(f, view_model.data(), view_model.size() * 4)) break;
        if (!read_names(f, model_names, 3)) break;
        scene_offsets.resize((size_t)n_scenes * n_scales + 1); scene_folder.resize(n_scenes);
        if (!read_exact(f, scene_offsets.data(), scene_offsets.size() * 8)) break;
        if (!read_exact(f, scene_folder.data(), scene_folder.size() * 4)) break;
        if (!read_names(f, scene_names, n_scenes)) break;
        scales.resize(n_scales);
        if (!read_exact(f, scales.data(), scales.size() * 4)) break;
        if (!read_u8_as_float(f, model_desc, (size_t)n_model_rows * 128)) break;
        model_kpt.resize((size_t)n_model_rows * 2);
        if (!read_exact(f, model_kpt.data(), model_kpt.size() * 4)) break;
        if (!read_u8_as_float(f, scene_desc, (size_t)n_scene_rows * 128)) break;
        scene_kpt.resize((size_t)n_scene_rows * 2);
        if (!read_exact(f, scene_kpt.data(), scene_kpt.size() * 4)) break;
        ok = true;
    } while (false);
    fclose(f);
    if (!ok && err) *err = "malformed feature cache " + path;
    return ok;
}

}  // namespace cvghost
