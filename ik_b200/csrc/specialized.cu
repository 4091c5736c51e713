// Registry of specialised kernels (see specialized.hpp).  Each specialisation is one generated header
// (ik_b200/csrc/gen/<name>.cuh, emitted at build time by tools/gen_kernel.py from ik_b200/specs/<name>.json and the
// product's own URDF flattener) compiled in its own translation unit (spec_<name>.cu) so they build in parallel.
#include "specialized.hpp"

#include <cstdio>

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

namespace ikb {

extern const SpecializedKernel kSpecCassieFeetPelvis;
extern const SpecializedKernel kSpecManipulatorTool;
extern const SpecializedKernel kSpecHumanoidLimbs;
extern const SpecializedKernel kSpecCassieDemo;
extern const SpecializedKernel kSpecCassieDemoPosture;

namespace {
const SpecializedKernel *const kRegistry[] = {&kSpecCassieFeetPelvis, &kSpecManipulatorTool, &kSpecHumanoidLimbs, &kSpecCassieDemo, &kSpecCassieDemoPosture, nullptr};
}

const SpecializedKernel *const *specialized_registry() { return kRegistry; }

// ---- plugins: specialisations compiled AFTER the library was built (ik_b200/specialise.py: any URDF / task list -> the
// generator -> nvcc -> a shared object exporting `ikb_spec_plugin`) -------------------------------------------------
namespace {
std::mutex g_plugin_mu;
std::vector<const SpecializedKernel *> g_plugins;
std::vector<std::string> g_plugin_paths;
bool g_env_loaded = false;
}  // namespace

int load_specialisation_plugin(const char *path, std::string *err) {
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    for (const std::string &p : g_plugin_paths)
        if (p == path) return 0;   // already loaded
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        if (err) *err = std::string("dlopen: ") + dlerror();
        return 1;
    }
    using Fn = const SpecializedKernel *(*)();
    Fn fn = (Fn)dlsym(h, "ikb_spec_plugin");
    const SpecializedKernel *k = fn ? fn() : nullptr;
    if (!k || !k->name || !k->matches || !k->launch64 || !k->launch32) {
        if (err) *err = std::string(path) + " does not export a complete ikb_spec_plugin";
        dlclose(h);
        return 1;
    }
    g_plugins.push_back(k);
    g_plugin_paths.push_back(path);
    return 0;
}

static void load_env_plugins() {   // IKB_SPEC_PLUGINS=a.so:b.so, read once
    {
        std::lock_guard<std::mutex> lk(g_plugin_mu);
        if (g_env_loaded) return;
        g_env_loaded = true;
    }
    const char *e = std::getenv("IKB_SPEC_PLUGINS");
    if (!e) return;
    std::string all(e);
    size_t pos = 0;
    while (pos <= all.size()) {
        const size_t end = all.find(':', pos);
        const std::string one = all.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
        if (!one.empty()) {
            std::string err;
            if (load_specialisation_plugin(one.c_str(), &err)) std::fprintf(stderr, "[ikb200] IKB_SPEC_PLUGINS: %s\n", err.c_str());
        }
        if (end == std::string::npos) break;
        pos = end + 1;
    }
}

const SpecializedKernel *find_specialized(const HostProblem &hp) {
    // IKB_FORCE_GENERIC=1 pins the table-driven kernel (used by the parity tests to cover both paths)
    const char *force = std::getenv("IKB_FORCE_GENERIC");
    if (force && force[0] == '1') return nullptr;
    for (const SpecializedKernel *const *k = kRegistry; *k; ++k)
        if ((*k)->matches(hp)) return *k;
    load_env_plugins();
    std::lock_guard<std::mutex> lk(g_plugin_mu);
    for (const SpecializedKernel *k : g_plugins)
        if (k->matches(hp)) return k;
    return nullptr;
}

const SpecializedKernel *find_near_miss(const HostProblem &hp, std::string *why) {
    for (const SpecializedKernel *const *k = kRegistry; *k; ++k)
        if ((*k)->near_miss && (*k)->near_miss(hp, why)) return *k;
    return nullptr;
}

}  // namespace ikb
