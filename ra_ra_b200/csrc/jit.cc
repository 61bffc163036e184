// jit.cc - run-time specialisation of the ply kernels for programs that are not in the pre-compiled table.
//
// The reference fuses every expression at COMPILE time: `B += A*C + v` is one template instantiation of ply_ravel's loop
// (ra/ply.hh:58-68) with the op inlined (ra/expr.hh:661). A C-ABI library cannot see the caller's templates, so the device
// analogue is done when the expression first arrives: the planner's canonical typed postfix program (static_progs.h, SProg)
// is handed to NVRTC as a constexpr (RACU_JIT_PROG) together with the SAME hand-written kernel bodies the pre-compiled table
// uses (static_device.h, embedded in the library as text), compiled for sm_100a with -fmad=false, and cached: in memory per
// device, and as a cubin on disk keyed by program + variant + a hash of the embedded sources. Same ops, same order, same
// types as the interpreter and the oracle, so results stay bit-identical; only the instruction stream is specialised.
//
// NVRTC is bound with dlopen at first use. If it is not there the planner keeps the interpreted kernels (ply_kernels.cu) -
// still the GPU, only slower; there is no CPU path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "launch.h"
#include "special.h"

namespace racu {

struct JitSource { char const * name; char const * text; };
static JitSource const jit_sources[] = {
#include "build/jit_sources.inc"
};

// ---- NVRTC, bound lazily
using nvrtcProgram = struct _nvrtcProgram *;
static struct Nvrtc
{
    int (*CreateProgram)(nvrtcProgram *, char const *, char const *, int, char const * const *, char const * const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, char const * const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    bool ok = false;
} nvrtc;

static bool
nvrtc_load()
{
    static std::once_flag once;
    std::call_once(once, []{
        void * h = nullptr;
        for (char const * name : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) { if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break; }
        if (!h) return;
        bool all = true;
#define RACU_SYM(NAME) do { nvrtc.NAME = reinterpret_cast<decltype(nvrtc.NAME)>(dlsym(h, "nvrtc" #NAME)); all = all && nvrtc.NAME; } while (0)
        RACU_SYM(CreateProgram); RACU_SYM(CompileProgram); RACU_SYM(GetCUBINSize); RACU_SYM(GetCUBIN); RACU_SYM(GetProgramLogSize);
        RACU_SYM(GetProgramLog); RACU_SYM(DestroyProgram);
#undef RACU_SYM
        nvrtc.ok = all;
    });
    return nvrtc.ok;
}

// ---- driver entry points (module load / launch), through the runtime like stencil.cu does for the tensor-map encoder
static struct Drv
{
    CUresult (*ModuleLoadData)(CUmodule *, void const *) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, char const *) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
    CUresult (*Occupancy)(int *, CUfunction, int, size_t) = nullptr;
    bool ok = false;
} drv;

static bool
drv_load()
{
    static std::once_flag once;
    std::call_once(once, []{
        auto get = [](char const * name) -> void * {
            void * p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaSuccess==cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) && q==cudaDriverEntryPointSuccess) return p;
            cudaGetLastError();
            return nullptr;
        };
        drv.ModuleLoadData = reinterpret_cast<decltype(drv.ModuleLoadData)>(get("cuModuleLoadData"));
        drv.ModuleGetFunction = reinterpret_cast<decltype(drv.ModuleGetFunction)>(get("cuModuleGetFunction"));
        drv.LaunchKernel = reinterpret_cast<decltype(drv.LaunchKernel)>(get("cuLaunchKernel"));
        drv.Occupancy = reinterpret_cast<decltype(drv.Occupancy)>(get("cuOccupancyMaxActiveBlocksPerMultiprocessor"));
        drv.ok = drv.ModuleLoadData && drv.ModuleGetFunction && drv.LaunchKernel;
    });
    return drv.ok;
}

// ---- keys and source text

static uint64_t
fnv1a(void const * data, size_t n, uint64_t h = 1469598103934665603ull)
{
    unsigned char const * p = static_cast<unsigned char const *>(data);
    for (size_t i=0; i<n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static uint64_t
sources_hash()
{
    static uint64_t h = []{
        uint64_t x = fnv1a("racu-jit-1", 10);
        for (JitSource const & s : jit_sources) x = fnv1a(s.text, std::strlen(s.text), x);
        return x;
    }();
    return h;
}

static int
fold_class(int fold_op)
{
    switch (fold_op) {
    case RACU_ASSIGN_ADD: case RACU_ASSIGN_SUB: return RACU_ASSIGN_ADD;
    case RACU_ASSIGN_AMIN: return RACU_ASSIGN_AMIN;
    case RACU_ASSIGN_AMAX: return RACU_ASSIGN_AMAX;
    default: return RACU_ASSIGN_MUL;
    }
}

std::string
jit_prog_text(SProg const & P)
{
    std::string s = "SProg {" + std::to_string(P.n) + ", {";
    for (int i=0; i<P.n; ++i) {
        SIns const & in = P.ins[i];
        s += "SIns {" + std::to_string(in.op) + ", " + std::to_string(in.arg) + ", " + std::to_string(in.type) + ", " + std::to_string(in.aux) + "}";
        if (i+1<P.n) s += ", ";
    }
    s += "}, " + std::to_string(P.nin) + ", " + std::to_string(P.yrole) + ", " + (P.fold ? "true" : "false") + ", {";
    for (int r=0; r<P.nin; ++r) { s += std::to_string(P.rtype[r]); if (r+1<P.nin) s += ", "; }
    s += "}, " + std::to_string(P.otype) + "}";
    return s;
}

// What, beyond the program, a specialisation is compiled for. Maps: GENERAL. Folds: the operand kind of every role and the re-read
// mask (compile-time in the fold bodies, static_device.h), and for fold-inner whether a warp or a CTA owns a kept row.
struct JitVariant { int general = 0, rows = 0, rmask = 0, lanes = 0; uint64_t kinds = 0; };

std::string
jit_kernel_text(SProg const & P, int mode, JitVariant const & v, int fold_op)
{
    std::string s = "#define RACU_JIT_PROG " + jit_prog_text(P) + "\n#include \"static_device.h\"\n";
    s += "extern \"C\" __global__ void __launch_bounds__(racu::KBLOCK)\nracu_jit_kernel(const __grid_constant__ racu::SParams p";
    std::string const fold = std::to_string(fold_class(fold_op)), kinds = std::to_string(v.kinds) + "ull";
    if (mode==K_MAP && v.lanes) s += ")\n{ racu::smap_lanes_body<racu::SP_JIT, " + kinds + ">(p); }\n";
    else if (mode==K_MAP) s += std::string(")\n{ racu::smap_body<racu::SP_JIT, ") + (v.general ? "true" : "false") + ", " + kinds + ">(p); }\n";
    else if (mode==K_FOLD_INNER) s += std::string(", const __grid_constant__ racu::FinishParams f)\n{ racu::") + (v.rows ? "sfold_inner_rows_body" : "sfold_inner_body")
                                      + "<racu::SP_JIT, " + fold + ", " + std::to_string(v.rmask) + ", " + kinds + ">(p, f); }\n";
    else s += ", const __grid_constant__ racu::FinishParams f)\n{ racu::sfold_outer_body<racu::SP_JIT, " + fold + ", " + kinds + ">(p, f); }\n";
    return s;
}

static std::string
cache_dir()
{
    if (char const * e = std::getenv("RACU_JIT_CACHE")) return e;
    Dl_info info;
    if (dladdr(reinterpret_cast<void *>(&sources_hash), &info) && info.dli_fname) { // next to libracu.so
        std::string p = info.dli_fname;
        size_t k = p.rfind('/');
        return (k==std::string::npos ? std::string(".") : p.substr(0, k)) + "/jit_cache";
    }
    return "jit_cache";
}

// Cache file = header + cubin. The header ties the file to the exact kernel text it was compiled from and to its own content, so
// a truncated, stale or foreign file is never handed to cuModuleLoadData.
struct CacheHeader { char magic[8]; uint64_t key; uint64_t size; uint64_t content_hash; };
static constexpr char CACHE_MAGIC[8] = {'R', 'A', 'C', 'U', 'J', 'I', 'T', '2'};

static bool
read_cache(std::string const & path, uint64_t key, std::vector<char> & out)
{
    FILE * f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    CacheHeader h;
    bool ok = std::fread(&h, 1, sizeof(h), f)==sizeof(h) && 0==std::memcmp(h.magic, CACHE_MAGIC, 8) && h.key==key && h.size>0 && h.size<(uint64_t(1)<<28);
    if (ok) {
        out.resize(size_t(h.size));
        ok = std::fread(out.data(), 1, out.size(), f)==out.size() && EOF==std::fgetc(f) && fnv1a(out.data(), out.size())==h.content_hash;
    }
    std::fclose(f);
    if (!ok) { out.clear(); std::remove(path.c_str()); } // recompiled (and rewritten) by the caller
    return ok;
}

static void
write_cache(std::string const & dir, std::string const & path, uint64_t key, std::vector<char> const & data)
{
    mkdir(dir.c_str(), 0755);
    std::string tmp = path + ".tmp" + std::to_string(long(getpid()));
    FILE * f = std::fopen(tmp.c_str(), "wb");
    if (!f) return;
    CacheHeader h;
    std::memcpy(h.magic, CACHE_MAGIC, 8);
    h.key = key; h.size = data.size(); h.content_hash = fnv1a(data.data(), data.size());
    bool ok = std::fwrite(&h, 1, sizeof(h), f)==sizeof(h) && std::fwrite(data.data(), 1, data.size(), f)==data.size();
    ok = (0==std::fclose(f)) && ok;
    if (ok) std::rename(tmp.c_str(), path.c_str()); else std::remove(tmp.c_str());
}

static int64_t jit_compiles = 0; // NVRTC compilations in this process (disk cache hits do not count)

// Compile (or fetch from the disk cache) the cubin of one specialisation. Needs no GPU. force: ignore (and replace) the cached file.
static bool
jit_cubin(SProg const & P, int mode, JitVariant const & variant, int fold_op, std::vector<char> & cubin, std::string & err, bool force=false)
{
    std::string const src = jit_kernel_text(P, mode, variant, fold_op);
    uint64_t const key = fnv1a(src.data(), src.size(), sources_hash());
    char name[40];
    std::snprintf(name, sizeof(name), "%016llx.cubin", static_cast<unsigned long long>(key));
    std::string const dir = cache_dir(), path = dir + "/" + name;
    bool const use_disk = !std::getenv("RACU_JIT_NO_DISK");
    if (force) std::remove(path.c_str());
    else if (use_disk && read_cache(path, key, cubin)) return true;
    if (!nvrtc_load()) { err = "NVRTC (libnvrtc.so.12) is not available"; return false; }
    constexpr int nh = int(sizeof(jit_sources)/sizeof(jit_sources[0]));
    char const * names[nh]; char const * texts[nh];
    for (int i=0; i<nh; ++i) { names[i] = jit_sources[i].name; texts[i] = jit_sources[i].text; }
    nvrtcProgram prog = nullptr;
    if (nvrtc.CreateProgram(&prog, src.c_str(), "racu_jit.cu", nh, texts, names)) { err = "nvrtcCreateProgram failed"; return false; }
    char const * opts[] = {"--gpu-architecture=sm_100a", "-std=c++20", "-fmad=false", "-default-device", "-lineinfo"};
    int const rc = nvrtc.CompileProgram(prog, int(sizeof(opts)/sizeof(opts[0])), opts);
    if (rc) {
        size_t n = 0;
        nvrtc.GetProgramLogSize(prog, &n);
        std::string log(n, ' ');
        if (n) nvrtc.GetProgramLog(prog, log.data());
        err = "nvrtcCompileProgram failed:\n" + log.substr(0, 4000);
        nvrtc.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    if (nvrtc.GetCUBINSize(prog, &n) || 0==n) { err = "nvrtcGetCUBINSize failed"; nvrtc.DestroyProgram(&prog); return false; }
    cubin.resize(n);
    int const r2 = nvrtc.GetCUBIN(prog, cubin.data());
    nvrtc.DestroyProgram(&prog);
    if (r2) { err = "nvrtcGetCUBIN failed"; return false; }
    ++jit_compiles;
    if (use_disk) write_cache(dir, path, key, cubin);
    return true;
}

// ---- per-device kernel cache

static std::mutex jit_mutex;
static std::map<std::string, CUfunction> jit_loaded;   // key: device, mode, variant, fold class + the program's bytes

static JitVariant
variant_of(Plan const & plan)
{
    SParams const & sp = plan.sp;
    JitVariant v;
    for (int k=0; k<plan.jit_prog.nin; ++k) { v.kinds |= uint64_t(sp.kind[k])<<(4*k); if (sp.mode==K_FOLD_INNER && sp.reused[k]) v.rmask |= 1<<k; }
    if (sp.mode==K_MAP) { v.general = sp.general ? 1 : 0; v.lanes = sp.lanes ? 1 : 0; return v; }
    v.rows = sp.mode==K_FOLD_INNER && sp.inner_rows;
    return v;
}

// Planner hook (plan.cc: g_jit_prepare): make sure the specialisation of plan.jit_prog exists. With a device, the kernel is
// loaded and plan.jit_fn set; without one (racu_prepare on a build host) the cubin is only compiled into the disk cache.
static bool
jit_prepare(Plan & plan, std::string & err)
{
    JitVariant const variant = variant_of(plan);
    int dev = -1;
    bool have_dev = cudaSuccess==cudaGetDevice(&dev) && dev>=0; // (cheap: no context work on the per-call path)
    if (!have_dev) cudaGetLastError();
    // in-memory key: no kernel text is generated on the per-call path
    SProg const & P = plan.jit_prog;
    int32_t const head[10] = {dev, plan.sp.mode, variant.general, variant.rows, variant.rmask, variant.lanes, int32_t(uint32_t(variant.kinds)), int32_t(uint32_t(variant.kinds>>32)),
                              plan.sp.mode==K_MAP ? 0 : fold_class(plan.sp.fold_op), P.n};
    std::string key(reinterpret_cast<char const *>(head), sizeof(head));
    key.append(reinterpret_cast<char const *>(P.ins), sizeof(SIns)*size_t(P.n));
    int32_t const tail[3] = {P.nin, P.yrole, int32_t(P.fold) | (int32_t(P.otype)<<8)};
    key.append(reinterpret_cast<char const *>(tail), sizeof(tail));
    key.append(reinterpret_cast<char const *>(P.rtype), size_t(P.nin));
    std::lock_guard<std::mutex> lock(jit_mutex);
    auto it = jit_loaded.find(key);
    if (it!=jit_loaded.end()) { plan.jit_fn = it->second; return true; }
    std::vector<char> cubin;
    if (!jit_cubin(plan.jit_prog, plan.sp.mode, variant, plan.sp.fold_op, cubin, err)) return false;
    if (have_dev && cudaSuccess!=cudaFree(nullptr)) { cudaGetLastError(); have_dev = false; } // make the primary context current before loading
    if (std::getenv("RACU_JIT_VERBOSE")) std::fprintf(stderr, "racu jit: specialised %s (mode %d general %d lanes %d rows %d rmask %d kinds %llx fold %d) for device %d, %zu bytes\n",
                                                      jit_prog_text(plan.jit_prog).c_str(), plan.sp.mode, variant.general, variant.lanes, variant.rows, variant.rmask, (unsigned long long)variant.kinds, plan.sp.fold_op, dev, cubin.size());
    if (!have_dev) { plan.jit_fn = nullptr; return true; }
    if (!drv_load()) { err = "driver entry points for module loading are not available"; return false; }
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    if (CUresult r = drv.ModuleLoadData(&mod, cubin.data())) {
        // a cached cubin the driver will not take (another toolkit, a damaged file that still hashes): compile it again, once
        if (!jit_cubin(plan.jit_prog, plan.sp.mode, variant, plan.sp.fold_op, cubin, err, true)) return false;
        if (CUresult r2 = drv.ModuleLoadData(&mod, cubin.data())) { err = "cuModuleLoadData failed: " + std::to_string(int(r)) + ", after recompiling: " + std::to_string(int(r2)); return false; }
    }
    if (CUresult r = drv.ModuleGetFunction(&fn, mod, "racu_jit_kernel")) { err = "cuModuleGetFunction failed: " + std::to_string(int(r)); return false; }
    jit_loaded.emplace(key, fn);
    plan.jit_fn = fn;
    return true;
}

cudaError_t
launch_jit(Plan const & plan, cudaStream_t stream)
{
    if (!plan.jit_fn || !drv_load()) return cudaErrorInvalidDeviceFunction;
    void * args[2] = {const_cast<SParams *>(&plan.sp), const_cast<FinishParams *>(&plan.fp)};
    CUresult r = drv.LaunchKernel(static_cast<CUfunction>(plan.jit_fn), plan.grid_x, plan.grid_y, 1, KBLOCK, 1, 1, 0, stream, args, nullptr);
    return cudaError_t(r); // driver and runtime share the error numbering (CUDA_SUCCESS == cudaSuccess == 0)
}

int
jit_occupancy(Plan const & plan)
{
    if (!plan.jit_fn || !drv_load() || !drv.Occupancy) return 0;
    static std::map<void *, int> cache;
    std::lock_guard<std::mutex> lock(jit_mutex);
    auto it = cache.find(plan.jit_fn);
    if (it!=cache.end()) return it->second;
    int n = 0;
    if (CUDA_SUCCESS!=drv.Occupancy(&n, static_cast<CUfunction>(plan.jit_fn), KBLOCK, 0)) return 0;
    cache.emplace(plan.jit_fn, n);
    return n;
}

int64_t jit_compile_count() { return jit_compiles; }

extern bool (*g_jit_prepare)(Plan &, std::string &);
static struct JitInstall { JitInstall() { g_jit_prepare = &jit_prepare; } } jit_install;

} // namespace racu
