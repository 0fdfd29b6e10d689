// eh_jit.cu -- NVRTC specialisation of traced process models (see eh_jit.h).  Host code only.
//
//   program (PmProgData)  --jit_source-->  C++ functor PmTraced + the three kernel instantiations
//                         --NVRTC (dlopen'ed; -arch=sm_100a, headers embedded at build time)-->  cubin (+ disk cache)
//                         --cudaLibraryLoadData / cudaLibraryGetKernel-->  kernel handles for the launch code of eh_lib.cu
//
// The functor evaluates exactly the formulas of the interpreter (PmProgram::eval / ::bwd, eh_pm.cuh), one C++ statement
// per instruction; what changes is that the compiler sees them: no switch, no value arrays in local memory, constants
// folded, dead adjoints (those of constants and forcings) removed, multiply-adds contracted.
#include "eh_jit.h"

#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "eh_pm.cuh"

#include "../_build/eh_jit_headers.inc"

namespace eh {

namespace {

const char* act_token(int act)
{
    switch (act) {
    case ACT_TANH: return "ACT_TANH";
    case ACT_SIGMOID: return "ACT_SIGMOID";
    case ACT_RELU: return "ACT_RELU";
    case ACT_SWISH: return "ACT_SWISH";
    default: return "ACT_IDENTITY";
    }
}

std::string flit(float x)
{
    char buf[64];
    if (x != x) return "__int_as_float(0x7fc00000)";
    if (x - x != 0.f) return x > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
    snprintf(buf, sizeof buf, "%af", (double)x);   // hexadecimal floating literal: exact
    return buf;
}

std::string vn(int i) { return "v" + std::to_string(i); }
std::string gn(int i) { return "g" + std::to_string(i); }

// the forward statement of instruction i (same formulas as PmProgram::eval)
std::string fwd_stmt(const PmProgData& pd, int i)
{
    const int op = pd.op[i];
    const std::string x = vn(pd.a[i]), y = vn(pd.b[i]);
    std::string r;
    switch (op) {
    case POP_CONST: r = flit(pd.imm[i]); break;
    case POP_FORCING: r = "f[" + std::to_string(pd.a[i]) + "]"; break;
    case POP_PARAM: r = "p[" + std::to_string(pd.a[i]) + "]"; break;
    case POP_ADD: r = x + " + " + y; break;
    case POP_SUB: r = x + " - " + y; break;
    case POP_MUL: r = x + " * " + y; break;
    case POP_DIV: r = x + " / " + y; break;
    case POP_POW: r = "powf(" + x + ", " + y + ")"; break;
    case POP_MIN: r = x + " < " + y + " ? " + x + " : " + y; break;
    case POP_MAX: r = x + " > " + y + " ? " + x + " : " + y; break;
    case POP_NEG: r = "-" + x; break;
    case POP_EXP: r = "expf(" + x + ")"; break;
    case POP_LOG: r = "logf(" + x + ")"; break;
    case POP_SQRT: r = "sqrtf(" + x + ")"; break;
    case POP_TANH: r = "tanhf(" + x + ")"; break;
    case POP_SIGMOID: r = "1.f / (1.f + expf(-" + x + "))"; break;
    case POP_ABS: r = "fabsf(" + x + ")"; break;
    case POP_SIN: r = "sinf(" + x + ")"; break;
    default: r = "cosf(" + x + ")"; break;
    }
    return "        const float " + vn(i) + " = " + r + ";\n";
}

// the reverse statements of instruction i (same formulas as PmProgram::bwd)
std::string bwd_stmt(const PmProgData& pd, int i)
{
    const int op = pd.op[i], ia = pd.a[i], ib = pd.b[i];
    if (op == POP_CONST || op == POP_FORCING) return "";
    const std::string gi = gn(i), vi = vn(i);
    if (op == POP_PARAM) return "        q" + std::to_string(ia) + " += " + gi + ";\n";
    const std::string x = vn(ia), y = vn(ib), ga = gn(ia), gb = gn(ib);
    std::string s;
    auto A = [&](const std::string& e) { s += "        " + ga + " += " + e + ";\n"; };
    auto B = [&](const std::string& e) { s += "        " + gb + " += " + e + ";\n"; };
    switch (op) {
    case POP_ADD: A(gi); B(gi); break;
    case POP_SUB: A(gi); B("-" + gi); break;
    case POP_MUL: A(gi + " * " + y); B(gi + " * " + x); break;
    case POP_DIV:
        s += "        { const float t = " + gi + " / " + y + "; " + ga + " += t; " + gb + " += -t * " + vi + "; }\n";
        break;
    case POP_POW:
        A(gi + " * " + y + " * powf(" + x + ", " + y + " - 1.f)");
        B(gi + " * " + vi + " * logf(" + x + ")");
        break;
    case POP_MIN: s += "        if (" + x + " < " + y + ") " + ga + " += " + gi + "; else " + gb + " += " + gi + ";\n"; break;
    case POP_MAX: s += "        if (" + x + " > " + y + ") " + ga + " += " + gi + "; else " + gb + " += " + gi + ";\n"; break;
    case POP_NEG: A("-" + gi); break;
    case POP_EXP: A(gi + " * " + vi); break;
    case POP_LOG: A(gi + " / " + x); break;
    case POP_SQRT: A(gi + " / (2.f * " + vi + ")"); break;
    case POP_TANH: A(gi + " * (1.f - " + vi + " * " + vi + ")"); break;
    case POP_SIGMOID: A(gi + " * " + vi + " * (1.f - " + vi + ")"); break;
    case POP_ABS: A(x + " < 0.f ? -" + gi + " : (" + x + " > 0.f ? " + gi + " : 0.f)"); break;
    case POP_SIN: A(gi + " * cosf(" + x + ")"); break;
    default: A("-" + gi + " * sinf(" + x + ")"); break;
    }
    return s;
}

uint64_t fnv1a(uint64_t h, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// ---- NVRTC through dlopen: the library has no link-time dependency on it ----
struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    nvrtcResult (*Version)(int*, int*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    std::string why;
};

const Nvrtc& nvrtc()
{
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("EH_NVRTC_LIB");
        const char* cand[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* c : cand) {
            if (!c || !*c) continue;
            n.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) { n.why = "libnvrtc.so.12 not found (set EH_NVRTC_LIB)"; return; }
#define EH_SYM(field, sym)                                              \
    *(void**)(&n.field) = dlsym(n.h, sym);                              \
    if (!n.field) { n.why = std::string("symbol missing in libnvrtc: ") + sym; n.h = nullptr; return; }
        EH_SYM(CreateProgram, "nvrtcCreateProgram")
        EH_SYM(DestroyProgram, "nvrtcDestroyProgram")
        EH_SYM(CompileProgram, "nvrtcCompileProgram")
        EH_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        EH_SYM(GetCUBIN, "nvrtcGetCUBIN")
        EH_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        EH_SYM(GetProgramLog, "nvrtcGetProgramLog")
        EH_SYM(AddNameExpression, "nvrtcAddNameExpression")
        EH_SYM(GetLoweredName, "nvrtcGetLoweredName")
        EH_SYM(Version, "nvrtcVersion")
        EH_SYM(GetErrorString, "nvrtcGetErrorString")
#undef EH_SYM
    });
    return n;
}

// headers the kernel sources name that NVRTC has no copy of
const char* const k_cstdint =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long long uintptr_t;\n";

std::string cache_dir()
{
    const char* e = getenv("EH_JIT_CACHE");
    if (e && *e) return e;
    const char* x = getenv("XDG_CACHE_HOME");
    if (x && *x) return std::string(x) + "/easyhybrid_b200";
    const char* h = getenv("HOME");
    if (h && *h) return std::string(h) + "/.cache/easyhybrid_b200";
    return "/tmp/easyhybrid_b200_jit";
}

void mkdirs(const std::string& d)
{
    for (size_t i = 1; i <= d.size(); i++)
        if (i == d.size() || d[i] == '/') mkdir(d.substr(0, i).c_str(), 0755);
}

// cache file: "EHJIT1\n" name0 "\n" name1 "\n" name2 "\n" <decimal size> "\n" <cubin bytes>
bool cache_read(const std::string& path, std::string* cubin, std::string names[3])
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char line[1024];
    bool ok = fgets(line, sizeof line, f) && !strcmp(line, "EHJIT1\n");
    for (int i = 0; ok && i < 3; i++) {
        ok = fgets(line, sizeof line, f) != nullptr;
        if (ok) { names[i] = line; if (!names[i].empty() && names[i].back() == '\n') names[i].pop_back(); ok = !names[i].empty(); }
    }
    size_t n = 0;
    if (ok) { ok = fgets(line, sizeof line, f) != nullptr; n = ok ? (size_t)strtoull(line, nullptr, 10) : 0; ok = ok && n > 0 && n < (64u << 20); }
    if (ok) { cubin->resize(n); ok = fread(&(*cubin)[0], 1, n, f) == n; }
    fclose(f);
    return ok;
}

void cache_write(const std::string& dir, const std::string& path, const std::string& cubin, const std::string names[3])
{
    mkdirs(dir);
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;   // a read-only cache directory only costs the next compilation
    fprintf(f, "EHJIT1\n%s\n%s\n%s\n%zu\n", names[0].c_str(), names[1].c_str(), names[2].c_str(), cubin.size());
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    if (fclose(f) != 0 || !ok || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
}

}  // namespace

std::string jit_source(const PmProgData& pd, const Variant& shape, const unsigned char* unit_act)
{
    std::string s;
    s += "// generated by eh_jit.cu: traced process model, " + std::to_string(pd.len) + " instructions\n";
    s += "#include \"eh_epoch_kernel.cuh\"\n#include \"eh_eval_kernel.cuh\"\nnamespace eh {\n";
    s += "struct PmTraced {\n";
    s += "    static constexpr int ID = PM_PROGRAM, NPS = PmProgram::NPS, NF = PmProgram::NF, NT = PmProgram::NT;\n";
    s += "    static constexpr bool DYNAMIC = true;\n";
    if (!unit_act) {
        s += "    static constexpr bool UNIT_ACT = false;\n";
    } else {
        // activation of unit j of hidden layer l as a chain of comparisons over runs of equal codes: plain arithmetic on
        // what are constants after unrolling
        s += "    static constexpr bool UNIT_ACT = true;\n";
        s += "    __device__ __forceinline__ static constexpr int unit_act(int l, int j)\n    {\n";
        for (int l = 1; l <= 3; l++) {
            s += "        if (l == " + std::to_string(l) + ") return ";
            const unsigned char* row = unit_act + (l - 1) * 32;
            std::string e;
            int j = 0;
            while (j < 32) {
                int k = j;
                while (k < 32 && row[k] == row[j]) k++;
                if (k < 32) e += "j < " + std::to_string(k) + " ? " + std::to_string((int)row[j]) + " : ";
                else e += std::to_string((int)row[j]);
                j = k;
            }
            s += e + ";\n";
        }
        s += "        return 0;\n    }\n";
    }
    s += "    static constexpr int NSV = " + std::to_string(pd.len > 0 ? pd.len : 1) + ";\n";
    s += "    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx&, float* y, float* sv)\n    {\n";
    for (int i = 0; i < pd.len; i++) s += fwd_stmt(pd, i);
    for (int t = 0; t < PmProgram::NT; t++)
        s += "        y[" + std::to_string(t) + "] = " + (t < pd.nt ? vn(pd.out[t]) : std::string("0.f")) + ";\n";
    for (int i = 0; i < pd.len; i++) s += "        sv[" + std::to_string(i) + "] = " + vn(i) + ";\n";
    s += "    }\n";
    s += "    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx&, const float* y,\n"
         "                                               const float* sv, const float* gy, float* gp)\n    {\n";
    for (int i = 0; i < pd.len; i++) s += "        const float " + vn(i) + " = sv[" + std::to_string(i) + "];\n";
    for (int i = 0; i < pd.len; i++) s += "        float " + gn(i) + " = 0.f;\n";
    for (int q = 0; q < MAXPS; q++) s += "        float q" + std::to_string(q) + " = 0.f;\n";
    for (int t = 0; t < PmProgram::NT && t < pd.nt; t++) s += "        " + gn(pd.out[t]) + " += gy[" + std::to_string(t) + "];\n";
    for (int i = pd.len - 1; i >= 0; i--) s += bwd_stmt(pd, i);
    for (int q = 0; q < MAXPS; q++) s += "        gp[" + std::to_string(q) + "] = q" + std::to_string(q) + ";\n";
    s += "        (void)p; (void)f; (void)y;";
    for (int i = 0; i < pd.len; i++) s += " (void)" + vn(i) + "; (void)" + gn(i) + ";";
    s += "\n    }\n};\n";
    char cfg[256];
    snprintf(cfg, sizeof cfg, "using CfgJ = StepCfg<%d, %d, %d, %d, %s, true, PmTraced>;\nusing EngJ = EngFfma<CfgJ>;\n", shape.P, shape.NH,
             shape.H, shape.NOUT, act_token(shape.act));
    s += cfg;
    s += "}  // namespace eh\n";
    return s;
}

bool jit_compile(const PmProgData& pd, const Variant& shape, const unsigned char* unit_act, std::string* cubin, std::string names[3],
                 std::string* tag, bool* from_cache, double* seconds, std::string* err)
{
    if (shape.pm != PM_PROGRAM || shape.engine != 0) { *err = "run-time specialisation starts from a generic FFMA2 variant"; return false; }
    for (int i = 0; i < pd.len; i++) {
        const int op = pd.op[i];
        const bool leaf = op == POP_CONST || op == POP_FORCING || op == POP_PARAM, binary = op >= POP_ADD && op <= POP_MAX;
        if (!leaf && (pd.a[i] < 0 || pd.a[i] >= i || (binary && (pd.b[i] < 0 || pd.b[i] >= i)))) { *err = "malformed program"; return false; }
        if (op == POP_PARAM && (pd.a[i] < 0 || pd.a[i] >= MAXPS)) { *err = "malformed program"; return false; }
        if (op == POP_FORCING && (pd.a[i] < 0 || pd.a[i] >= PmProgram::NF)) { *err = "malformed program"; return false; }
    }
    const std::string src = jit_source(pd, shape, unit_act);
    const Nvrtc& n = nvrtc();
    int vmaj = 0, vmin = 0;
    if (n.h) n.Version(&vmaj, &vmin);
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, src.data(), src.size());
    for (int i = 0; i < eh_jit_num_headers; i++) h = fnv1a(h, eh_jit_header_sources[i], strlen(eh_jit_header_sources[i]));
    h = fnv1a(h, "sm_100a", 7);
    char hs[32];
    snprintf(hs, sizeof hs, "%016llx", (unsigned long long)h);
    char tg[160];
    snprintf(tg, sizeof tg, "nvrtc/PmTraced#%.8s/P%d/NH%d/H%d/O%d/%s", hs, shape.P, shape.NH, shape.H, shape.NOUT,
             unit_act ? "ACT_PER_UNIT" : act_token(shape.act));
    *tag = tg;
    *seconds = 0.0;
    const std::string dir = cache_dir(), path = dir + "/" + hs + ".ehjit";
    const bool use_cache = !getenv("EH_JIT_NO_CACHE");
    if (use_cache && cache_read(path, cubin, names)) { *from_cache = true; return true; }
    *from_cache = false;
    if (!n.h) { *err = "NVRTC unavailable: " + n.why; return false; }

    std::vector<const char*> hn(eh_jit_header_names, eh_jit_header_names + eh_jit_num_headers), hsrc(eh_jit_header_sources, eh_jit_header_sources + eh_jit_num_headers);
    const char* stubs[] = {"cstdint", "cstdio", "cstdlib", "cmath", "cuda_runtime.h"};
    for (const char* st : stubs) { hn.push_back(st); hsrc.push_back(!strcmp(st, "cstdint") ? k_cstdint : "#pragma once\n"); }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.CreateProgram(&prog, src.c_str(), "eh_traced.cu", (int)hn.size(), hsrc.data(), hn.data());
    if (r != NVRTC_SUCCESS) { *err = std::string("nvrtcCreateProgram: ") + n.GetErrorString(r); return false; }
    const char* exprs[3] = {"&eh::k_step<eh::EngJ>", "&eh::k_epoch<eh::EngJ>", "&eh::k_eval<eh::CfgJ>"};
    for (const char* e : exprs) n.AddNameExpression(prog, e);
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "-lineinfo"};
    const auto t0 = std::chrono::steady_clock::now();
    r = n.CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        if (log.size() > 1500) log.resize(1500);
        *err = std::string("NVRTC: ") + n.GetErrorString(r) + "\n" + log;
        n.DestroyProgram(&prog);
        return false;
    }
    bool ok = true;
    for (int i = 0; i < 3 && ok; i++) {
        const char* low = nullptr;
        ok = n.GetLoweredName(prog, exprs[i], &low) == NVRTC_SUCCESS && low;
        if (ok) names[i] = low;
    }
    size_t cs = 0;
    ok = ok && n.GetCUBINSize(prog, &cs) == NVRTC_SUCCESS && cs > 0;
    if (ok) { cubin->resize(cs); ok = n.GetCUBIN(prog, &(*cubin)[0]) == NVRTC_SUCCESS; }
    n.DestroyProgram(&prog);
    if (!ok) { *err = "NVRTC produced no cubin / kernel names"; return false; }
    if (use_cache) cache_write(dir, path, *cubin, names);
    return true;
}

bool jit_load(const std::string& cubin, const std::string names[3], JitKernels* out, std::string* err)
{
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { *err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); cudaGetLastError(); return false; }
    const void** dst[3] = {&out->k_step, &out->k_epoch, &out->k_eval};
    for (int i = 0; i < 3; i++) {
        cudaKernel_t k = nullptr;
        e = cudaLibraryGetKernel(&k, lib, names[i].c_str());
        if (e != cudaSuccess) {
            *err = "cudaLibraryGetKernel(" + names[i] + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            cudaLibraryUnload(lib);
            return false;
        }
        *dst[i] = (const void*)k;
    }
    out->lib = lib;
    return true;
}

void jit_unload(JitKernels* k)
{
    if (k && k->lib) { cudaLibraryUnload(k->lib); k->lib = nullptr; }
}

cudaError_t jit_prepare(const JitKernels& k, size_t step_smem, size_t eval_smem)
{
    cudaError_t e = cudaFuncSetAttribute(k.k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k.k_epoch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k.k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_smem);
}

cudaError_t jit_launch(const void* kernel, const void* args, int grid, int threads, size_t smem, cudaStream_t st, bool pdl)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    void* argv[] = {const_cast<void*>(args)};
    return cudaLaunchKernelExC(&cfg, kernel, argv);
}

cudaError_t jit_launch_cooperative(const void* kernel, const void* args, int grid, int threads, size_t smem, cudaStream_t st)
{
    if (getenv("EH_NO_COOP")) return jit_launch(kernel, args, grid, threads, smem, st, false);
    void* argv[] = {const_cast<void*>(args)};
    return cudaLaunchCooperativeKernel(kernel, dim3((unsigned)grid), dim3((unsigned)threads), argv, smem, st);
}

cudaError_t jit_max_grid(const void* kernel, int threads, size_t smem, int* max_ctas)
{
    int per_sm = 0, dev = 0, nsm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    *max_ctas = per_sm * nsm;
    return cudaSuccess;
}

}  // namespace eh
