// lcu_internal.h -- shared declarations of the C-ABI CUDA layer.
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "lensed_cuda.h"

namespace lcu {

// error reporting -----------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

// program assembly ------------------------------------------------------------

struct ObjectInfo
{
    std::string name;       // file name without .cl
    std::string ident;      // name made safe for C identifiers
    std::string wrapped;    // plugin text inside its name-mangling macros + metadata export
    int type = 0;
    size_t bytes = 0;       // sizeof(struct data_<ident>)
    size_t words = 0;       // 4-byte words, rounded up (src/input/objects.c:139)
    std::vector<lcu_param> params;
    bool pairable = false;  // the per-ray functions also compile for two rays per thread (shim.cuh)
    std::string pair_log;   // why not, if not
};

struct ModelObject
{
    const ObjectInfo* info = nullptr;
    size_t d = 0;           // word offset of the data block
    size_t p = 0;           // index of the first parameter
    std::vector<int> ipp;   // image-plane-prior flag per parameter
};

struct ProgramOptions
{
    size_t width = 0, height = 0;
    int psf = 0;
    size_t psfw = 0, psfh = 0;
    size_t nq = 0;
    size_t maxb = 1;
    bool fast_math = false;
    bool obj_const = true;
};

std::string read_text_file(const std::string& path, bool* ok);
std::string rewrite_literals(const std::string& text);
std::string make_ident(const std::string& name);
std::string wrap_object(const std::string& name, const std::string& ident, const std::string& text);

// src/kernel.c:235-399 and :401-656 equivalents: CUDA text of
// lcu_compute() and lcu_set_params_body() for an object list
std::string generate_compute(const std::vector<ModelObject>& objs, bool pair = false);
std::string strip_for_pair(const std::string& text);
std::string generate_set_params(const std::vector<ModelObject>& objs);

// NVRTC: source (+ named headers) -> sm_100a cubin
typedef std::pair<std::string, std::string> Header;    // (include name, text)
bool compile_cubin(const std::string& source, const std::vector<Header>& headers,
                   const std::vector<std::string>& options,
                   std::vector<char>* cubin, std::string* log);

// read the initialised bytes of a global symbol out of a cubin (ELF64)
bool cubin_symbol(const std::vector<char>& cubin, const std::string& symbol,
                  const unsigned char** bytes, size_t* size);

// registers per thread and stack bytes of a kernel of a cubin (.nv.info records)
bool cubin_kernel_usage(const std::vector<char>& cubin, const std::string& kernel, unsigned* regs, unsigned* stack);

// quadrature -------------------------------------------------------------------
int quad_rule_count();
const char* quad_rule_name(int i);
const char* quad_rule_info(int i);
int quad_rule(const char* rule, double sx, double sy, float* qq, float* ww);

} // namespace lcu

struct lcu_ctx
{
    int device = -1;                    // < 0: compile-only
    std::string kernel_dir, objects_dir;
    std::string shim, object_hdr, kernels;      // kernel/*.cuh, lensed.cu text
    std::map<std::string, lcu::ObjectInfo> objects;
    std::map<std::string, std::vector<char>> cubins;            // compiled programs by options + source text
    int sm_count = 0;

    const lcu::ObjectInfo* object(const std::string& name);     // loads + compiles on first use
    std::vector<lcu::Header> headers() const;
    std::vector<std::string> build_options(unsigned flags) const;
    // NVRTC with an in-process cache: identical program text + options compile once
    bool compile(const std::string& source, unsigned flags, std::vector<char>* cubin, std::string* log);
};
