// lcu_program.cpp -- object loading, program assembly, NVRTC build, metadata.
//
// Replaces the reference's src/kernel.c (load_object :731-816, main_program
// :838-879, compute_kernel :235-399, set_params_kernel :401-656,
// kernel_options :881-944) and the metadata round trip of
// src/input/objects.c:72-239.  Same contract for object files; different
// mechanics: C++ instead of OpenCL C, NVRTC instead of clBuildProgram, and
// metadata read from the compiled module instead of from meta kernels.

#include "lcu_internal.h"

#include <nvrtc.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <regex>
#include <sstream>

namespace lcu {

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_error;

void set_error(const char* fmt, ...)
{
    char buf[1 << 16];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

const char* get_error() { return g_error.c_str(); }

// ---------------------------------------------------------------------------
// text utilities
// ---------------------------------------------------------------------------
std::string read_text_file(const std::string& path, bool* ok)
{
    std::ifstream f(path, std::ios::binary);
    if(!f)
    {
        *ok = false;
        return std::string();
    }
    std::ostringstream ss;
    ss << f.rdbuf();
    *ok = true;
    std::string text = ss.str();
    // a UTF-8 byte order mark some editors put first is not part of the program
    if(text.size() >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF)
        text.erase(0, 3);
    return text;
}

// OpenCL vector literals "(float2)(a, b)" are a cast applied to a comma
// expression in C++; rewrite them to constructor calls "float2(a, b)".
std::string rewrite_literals(const std::string& text)
{
    static const std::regex lit(R"(\(\s*(float2|float4|mat22)\s*\)\s*\()");
    return std::regex_replace(text, lit, "$1(");
}

// object names go into identifiers (type_<name>, data_<name>, ...); the
// reference pastes them verbatim (src/kernel.c:153-162), which cannot work for
// names like "sersic-old"; here every non-identifier character becomes '_'
std::string make_ident(const std::string& name)
{
    std::string id = name;
    for(char& c : id)
        if(!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_'))
            c = '_';
    return id;
}

// The pair copy of an object (two rays per thread, shim.cuh) needs only the
// functions that run per ray.  Everything that describes or fills the data
// block -- `type = ...;`, `params {...};`, `data {...};`, the set() function --
// and program-scope constants are blanked out of that copy (newlines kept, so
// that diagnostics still point at the right line): the data block has one
// layout, defined by the scalar copy, set() may branch on its arguments, which
// pairs cannot, and a constant is the same number for both rays.
// A small scanner rather than a parser: top-level items end at a ';' or at the
// '}' that closes a function body; comments, literals and preprocessor lines
// are skipped over.
std::string strip_for_pair(const std::string& text)
{
    std::string out = text;
    const size_t n = text.size();
    size_t item = 0;            // start of the current top-level item
    int depth = 0;
    bool body = false;          // the item has had a '{' ... '}' at depth 0
    auto is_id = [](char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_'; };
    // decide about the item [item, end) and start the next one
    auto close = [&](size_t end)
    {
        // identifiers outside comments up to the first '{' or '='
        std::vector<std::string> ids;
        bool call_set = false;
        bool variable = false;      // "<qualifiers> <type> name[...] = ...;": a program-scope constant
        for(size_t i = item; i < end; )
        {
            const char c = text[i];
            if(c == '/' && i + 1 < end && text[i + 1] == '/') { while(i < end && text[i] != '\n') ++i; continue; }
            if(c == '/' && i + 1 < end && text[i + 1] == '*') { i += 2; while(i + 1 < end && !(text[i] == '*' && text[i + 1] == '/')) ++i; i += 2; continue; }
            if(c == '=') { variable = true; break; }
            if(c == '{') break;
            if(is_id(c))
            {
                size_t j = i;
                while(j < end && is_id(text[j])) ++j;
                ids.push_back(text.substr(i, j - i));
                size_t k = j;
                while(k < end && (text[k] == ' ' || text[k] == '\t' || text[k] == '\n' || text[k] == '\r')) ++k;
                if(ids.back() == "set" && k < end && text[k] == '(')
                    call_set = true;
                i = j;
                if(c == '(') break;
                continue;
            }
            if(c == '(') { break; }
            ++i;
        }
        // program-scope constants stay what the scalar copy made of them (plain floats,
        // visible here through the using-directive of the pair namespace)
        const bool drop = call_set || variable || (!ids.empty() && (ids[0] == "type" || ids[0] == "params" || ids[0] == "data"));
        if(drop)
            for(size_t i = item; i < end; ++i)
                if(out[i] != '\n')
                    out[i] = ' ';
        item = end;
        body = false;
    };
    for(size_t i = 0; i < n; )
    {
        const char c = text[i];
        if(c == '/' && i + 1 < n && text[i + 1] == '/') { while(i < n && text[i] != '\n') ++i; continue; }
        if(c == '/' && i + 1 < n && text[i + 1] == '*') { i += 2; while(i + 1 < n && !(text[i] == '*' && text[i + 1] == '/')) ++i; i += 2; continue; }
        if(c == '"' || c == '\'')
        {
            const char q = c;
            ++i;
            while(i < n && text[i] != q) { if(text[i] == '\\') ++i; ++i; }
            ++i;
            continue;
        }
        if(c == '#' && depth == 0)
        {
            // preprocessor line (with continuations): an item of its own, kept
            size_t j = i;
            while(j < n && text[j] != '\n') { if(text[j] == '\\' && j + 1 < n && text[j + 1] == '\n') ++j; ++j; }
            item = j;
            i = j;
            continue;
        }
        if(c == '{') { ++depth; ++i; continue; }
        if(c == '}')
        {
            --depth;
            ++i;
            if(depth == 0)
                body = true;
            continue;
        }
        if(depth == 0 && body && c != ' ' && c != '\t' && c != '\n' && c != '\r' && c != ';')
        {
            // something new starts after a closed body: it was a function
            close(i);
            continue;
        }
        if(c == ';' && depth == 0) { ++i; close(i); continue; }
        ++i;
    }
    if(body)
        close(n);
    return out;
}

// The name-mangling wrapper of src/kernel.c:153-172 (OBJHEAD / OBJFOOT), plus
// the metadata export that replaces the meta_<name> / params_<name> kernels
// of src/kernel.c:41-62: type, sizeof(data) and the parameter count become an
// initialised device array the host reads back from the module image.
std::string wrap_object(const std::string& name, const std::string& id, const std::string& text)
{
    std::ostringstream s;
    s << "//----------------------------------------------------------------------------\n"
      << "// objects/" << name << ".cl\n"
      << "//----------------------------------------------------------------------------\n"
      // every copy of the plugin text lives in a namespace of its own (lcu_ray,
      // lcu_setter, lcu_pair): a helper function the plugin defines exists once
      // per copy, and argument-dependent lookup (the vector types are global)
      // must not find the other copies' versions
      << "namespace lcu_ray {\n"
      << "#define LCU_SHIM_ON\n#include \"shim.cuh\"\n"
      << "#if LCU_INTRINSICS_@KIND@\n#define LCU_INTRINSICS_ON\n#include \"shim.cuh\"\n#endif\n"
      << "#if LCU_ATANH_@KIND@\n#define LCU_ATANH_ON\n#include \"shim.cuh\"\n#endif\n"
      << "#define type const int type_" << id << "\n"
      << "#define params extern \"C\" __device__ const struct param lcu_parlst_" << id << "[] = \n"
      << "#define data struct data_" << id << "\n"
      << "#define deflection deflection_" << id << "\n"
      << "#define brightness brightness_" << id << "\n"
      << "#define foreground foreground_" << id << "\n"
      << "#define set set_" << id << "\n"
      << "#line 1 \"objects/" << name << ".cl\"\n"
      << rewrite_literals(text) << "\n"
      << "#undef type\n#undef params\n#undef data\n#undef deflection\n"
      << "#undef brightness\n#undef foreground\n#undef set\n"
      << "#if LCU_INTRINSICS_@KIND@\n#define LCU_INTRINSICS_OFF\n#include \"shim.cuh\"\n#endif\n"
      << "#if LCU_ATANH_@KIND@\n#define LCU_ATANH_OFF\n#include \"shim.cuh\"\n#endif\n"
      << "#define LCU_SHIM_OFF\n#include \"shim.cuh\"\n"
      << "extern \"C\" __device__ const unsigned int lcu_meta_" << id << "[3] = {\n"
      << "    (unsigned int)type_" << id << ",\n"
      << "    (unsigned int)sizeof(struct data_" << id << "),\n"
      << "    (unsigned int)(sizeof(lcu_parlst_" << id << ")/sizeof(struct param))\n"
      << "};\n"
      << "} // namespace lcu_ray\n\n";
    // Second copy of the object for the parameter setter (and the ray shots
    // of image-plane priors): same text, but its math built-ins are evaluated
    // in double and rounded once (shim.cuh: LCU_ACCURATE_ON).  set_params runs
    // one thread per parameter point, so this costs nothing measurable and
    // makes the object block agree with a correctly rounding host libm.
    s << "namespace lcu_setter {\n"
      << "#define LCU_SHIM_ON\n#include \"shim.cuh\"\n"
      << "#define LCU_ACCURATE_ON\n#include \"shim.cuh\"\n"
      << "#define type const int type_" << id << "\n"
      << "#define params const struct param parlst_" << id << "[] = \n"
      << "#define data struct data_" << id << "\n"
      << "#define deflection deflection_" << id << "\n"
      << "#define brightness brightness_" << id << "\n"
      << "#define foreground foreground_" << id << "\n"
      << "#define set set_" << id << "\n"
      << "#line 1 \"objects/" << name << ".cl\"\n"
      << rewrite_literals(text) << "\n"
      << "#undef type\n#undef params\n#undef data\n#undef deflection\n"
      << "#undef brightness\n#undef foreground\n#undef set\n"
      << "#define LCU_ACCURATE_OFF\n#include \"shim.cuh\"\n"
      << "#define LCU_SHIM_OFF\n#include \"shim.cuh\"\n"
      << "} // namespace lcu_setter\n\n";
    // Third copy, for the two-rays-per-thread render kernel: the per-ray
    // functions with float = a packed pair (shim.cuh), the data block shared
    // with the scalar copy.  Only compiled when every object of the model
    // can be typed that way (LCU_PAIR, decided in lcu_ctx::object).
    s << "#if LCU_PAIR\nnamespace lcu_pair {\nusing namespace ::lcu_ray;\n"
      << "#define LCU_SHIM_ON\n#include \"shim.cuh\"\n"
      << "#define LCU_PAIR_ON\n#include \"shim.cuh\"\n"
      << "#if LCU_INTRINSICS_@KIND@\n#define LCU_INTRINSICS_ON\n#include \"shim.cuh\"\n#endif\n"
      << "#if LCU_ATANH_@KIND@\n#define LCU_ATANH_ON\n#include \"shim.cuh\"\n#endif\n"
      << "#define data struct ::lcu_ray::data_" << id << "\n"
      << "#define deflection deflection_" << id << "\n"
      << "#define brightness brightness_" << id << "\n"
      << "#define foreground foreground_" << id << "\n"
      << "#line 1 \"objects/" << name << ".cl\"\n"
      << rewrite_literals(strip_for_pair(text)) << "\n"
      << "#undef data\n#undef deflection\n#undef brightness\n#undef foreground\n"
      << "#if LCU_INTRINSICS_@KIND@\n#define LCU_INTRINSICS_OFF\n#include \"shim.cuh\"\n#endif\n"
      << "#if LCU_ATANH_@KIND@\n#define LCU_ATANH_OFF\n#include \"shim.cuh\"\n#endif\n"
      << "#define LCU_PAIR_OFF\n#include \"shim.cuh\"\n"
      << "#define LCU_SHIM_OFF\n#include \"shim.cuh\"\n"
      << "} // namespace lcu_pair\n#endif\n\n";
    return s.str();
}

// ---------------------------------------------------------------------------
// generated device functions
// ---------------------------------------------------------------------------
static const char* DEFLECT =
    " -= dot(a, a) < HUGE_VALF ? a : lcu_float2(1E10f, 1E10f);\n";

// compute(): src/kernel.c:65-111 templates, :321-383 object loop.  Objects are
// visited in ini order.  A change of the (non-foreground) object type away
// from LENS closes the lens plane: the summed deflection is applied to the
// ray if finite.  Sources see the ray position y, foregrounds the image-plane
// position x.
std::string generate_compute(const std::vector<ModelObject>& objs, bool pair)
{
    // pair: the same function for two rays per thread, on the lcu_pair copies
    // of the objects (lcu_compute2, shim.cuh: packed pairs)
    const char* V2 = pair ? "lcu_pf2" : "lcu_float2";
    const char* NS = pair ? "lcu_pair::" : "lcu_ray::";
    const std::string DEFLECT = pair ? " -= lcu_pair_guard(a);\n" : lcu::DEFLECT;
    std::ostringstream s;
    s << "__device__ __forceinline__ " << (pair ? "lcu_pf lcu_compute2" : "float lcu_compute") << "(const uint* data, " << V2 << " x)\n{\n"
      << "    " << V2 << " y = x;\n"
      << "    " << (pair ? "lcu_pf" : "float") << " f = 0;\n";
    int type = 0, trigger = 0;
    bool open = false;
    // The reference starts every sum from zero ("float2 a = 0; a += ...",
    // "float f = 0; f += ...").  The first term is assigned instead: 0 + t == t
    // for every t except t = -0, whose sign no later operation can observe
    // (y - a, f + ...), and the compiler may not drop the addition itself.
    bool first_lens = false, first_light = true;
    for(const ModelObject& o : objs)
    {
        const int t = o.info->type;
        const std::string& id = o.info->ident;
        if(t != trigger && t != LCU_FOREGROUND)
        {
            if(trigger == LCU_LENS)
            {
                s << "        y" << DEFLECT << "    }\n";
                open = false;
            }
            trigger = t;
        }
        if(t != type)
        {
            if(t == LCU_LENS && !open)
            {
                s << "    {\n        " << V2 << " a;\n";
                open = true;
                first_lens = true;
            }
            type = t;
        }
        const char* ind = open ? "        " : "    ";
        if(t == LCU_LENS)
        {
            s << ind << (first_lens ? "a = " : "a += ") << NS << "deflection_" << id << "((struct lcu_ray::data_" << id << "*)(data + " << o.d << "), y);\n";
            first_lens = false;
        }
        else
        {
            s << ind << (first_light ? "f = " : "f += ") << NS << (t == LCU_SOURCE ? "brightness_" : "foreground_") << id
              << "((struct lcu_ray::data_" << id << "*)(data + " << o.d << "), " << (t == LCU_SOURCE ? "y" : "x") << ");\n";
            first_light = false;
        }
    }
    if(trigger == LCU_LENS)
        s << "        y" << DEFLECT << "    }\n";
    s << "    return f;\n}\n\n";
    return s.str();
}

// set_params(): src/kernel.c:114-150 templates, :455-633 object loop.  Each
// object's setter receives its parameters in declaration order; a position
// pair with image-plane priors is first shot through every lens in front of
// the current source plane (src/kernel.c:499-564).
std::string generate_set_params(const std::vector<ModelObject>& objs)
{
    std::ostringstream s;
    s << "__device__ __forceinline__ void lcu_set_params_body(uint* data, const float* params)\n{\n"
      << "    lcu_float2 x = 0;\n"
      << "    lcu_float2 a = 0;\n";
    int trigger = 0;
    size_t plane = 0;
    for(size_t i = 0; i < objs.size(); ++i)
    {
        const ModelObject& o = objs[i];
        const int t = o.info->type;
        const std::string& id = o.info->ident;
        if(t != trigger && t != LCU_FOREGROUND)
        {
            if(trigger == LCU_LENS && t == LCU_SOURCE)
                plane = i;
            trigger = t;
        }
        for(size_t j = 0; j < o.info->params.size(); ++j)
        {
            if(!o.ipp[j] || o.info->params[j].type != LCU_POSITION_X)
                continue;
            int trigger2 = 0;
            s << "    x = lcu_float2(params[" << (o.p + j) << "], params[" << (o.p + j + 1) << "]);\n";
            for(size_t k = 0; k < plane; ++k)
            {
                const ModelObject& l = objs[k];
                if(l.info->type != trigger2 && l.info->type != LCU_FOREGROUND)
                {
                    if(trigger2 == LCU_LENS)
                        s << "    x" << DEFLECT << "    a = 0;\n";
                    trigger2 = l.info->type;
                }
                if(l.info->type == LCU_LENS)
                    s << "    a += lcu_setter::deflection_" << l.info->ident << "((struct lcu_setter::data_" << l.info->ident
                      << "*)(data + " << l.d << "), x);\n";
            }
            if(trigger2 == LCU_LENS)
                s << "    x" << DEFLECT << "    a = 0;\n";
        }
        s << "    lcu_setter::set_" << id << "((struct lcu_setter::data_" << id << "*)(data + " << o.d << ")";
        for(size_t j = 0; j < o.info->params.size(); ++j)
        {
            if(o.ipp[j])
            {
                const int pt = o.info->params[j].type;
                s << ", " << (pt == LCU_POSITION_X ? "x.x" : pt == LCU_POSITION_Y ? "x.y" : "0");
            }
            else
                s << ", params[" << (o.p + j) << "]";
        }
        s << ");\n";
    }
    s << "    (void)x; (void)a;\n}\n\n";
    return s.str();
}

// ---------------------------------------------------------------------------
// NVRTC
// ---------------------------------------------------------------------------
bool compile_cubin(const std::string& source, const std::vector<Header>& headers,
                   const std::vector<std::string>& options,
                   std::vector<char>* cubin, std::string* log)
{
    nvrtcProgram prog = nullptr;
    std::vector<const char*> hsrc, hname;
    for(const Header& h : headers)
    {
        hname.push_back(h.first.c_str());
        hsrc.push_back(h.second.c_str());
    }
    nvrtcResult r = nvrtcCreateProgram(&prog, source.c_str(), "lensed_model.cu", (int)headers.size(),
                                       hsrc.data(), hname.data());
    if(r != NVRTC_SUCCESS)
    {
        *log = std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r);
        return false;
    }
    std::vector<const char*> opts;
    for(const std::string& o : options)
        opts.push_back(o.c_str());
    r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    if(n > 1)
    {
        log->resize(n);
        nvrtcGetProgramLog(prog, &(*log)[0]);
    }
    else
        log->clear();
    if(r != NVRTC_SUCCESS)
    {
        *log = std::string("nvrtcCompileProgram: ") + nvrtcGetErrorString(r) + "\n" + *log;
        nvrtcDestroyProgram(&prog);
        return false;
    }
    size_t size = 0;
    r = nvrtcGetCUBINSize(prog, &size);
    if(r != NVRTC_SUCCESS || size == 0)
    {
        *log += "\nnvrtcGetCUBINSize failed";
        nvrtcDestroyProgram(&prog);
        return false;
    }
    cubin->resize(size);
    nvrtcGetCUBIN(prog, cubin->data());
    nvrtcDestroyProgram(&prog);
    return true;
}

// ---------------------------------------------------------------------------
// cubin (ELF64) symbol reader
// ---------------------------------------------------------------------------
namespace {
struct Elf64Ehdr
{
    unsigned char ident[16];
    uint16_t type, machine;
    uint32_t version;
    uint64_t entry, phoff, shoff;
    uint32_t flags;
    uint16_t ehsize, phentsize, phnum, shentsize, shnum, shstrndx;
};
struct Elf64Shdr
{
    uint32_t name, type;
    uint64_t flags, addr, offset, size;
    uint32_t link, info;
    uint64_t addralign, entsize;
};
struct Elf64Sym
{
    uint32_t name;
    unsigned char info, other;
    uint16_t shndx;
    uint64_t value, size;
};
} // namespace

bool cubin_symbol(const std::vector<char>& cubin, const std::string& symbol,
                  const unsigned char** bytes, size_t* size)
{
    const unsigned char* base = reinterpret_cast<const unsigned char*>(cubin.data());
    const size_t len = cubin.size();
    if(len < sizeof(Elf64Ehdr) || memcmp(base, "\177ELF", 4) != 0 || base[4] != 2)
        return false;
    Elf64Ehdr eh;
    memcpy(&eh, base, sizeof(eh));
    if(eh.shoff == 0 || eh.shentsize != sizeof(Elf64Shdr) || eh.shoff + (uint64_t)eh.shnum*sizeof(Elf64Shdr) > len)
        return false;
    std::vector<Elf64Shdr> sh(eh.shnum);
    memcpy(sh.data(), base + eh.shoff, eh.shnum*sizeof(Elf64Shdr));
    for(const Elf64Shdr& st : sh)
    {
        if(st.type != 2 /* SHT_SYMTAB */ || st.link >= sh.size() || st.entsize != sizeof(Elf64Sym))
            continue;
        const Elf64Shdr& str = sh[st.link];
        if(st.offset + st.size > len || str.offset + str.size > len)
            return false;
        const size_t nsym = st.size/sizeof(Elf64Sym);
        for(size_t i = 0; i < nsym; ++i)
        {
            Elf64Sym sym;
            memcpy(&sym, base + st.offset + i*sizeof(Elf64Sym), sizeof(sym));
            if(sym.name >= str.size)
                continue;
            const char* nm = reinterpret_cast<const char*>(base + str.offset + sym.name);
            if(symbol != nm)
                continue;
            if(sym.shndx == 0 || sym.shndx >= sh.size())
                return false;
            const Elf64Shdr& sec = sh[sym.shndx];
            if(sec.type == 8 /* SHT_NOBITS */ || sym.value + sym.size > sec.size || sec.offset + sec.size > len)
                return false;
            *bytes = base + sec.offset + sym.value;
            *size = sym.size;
            return true;
        }
    }
    return false;
}

// Registers per thread and stack (spill) bytes of a kernel, from the
// EIATTR_REGCOUNT (0x2f) / EIATTR_MIN_STACK_SIZE (0x12) records of the cubin's
// .nv.info section; both carry (symbol index, value).
bool cubin_kernel_usage(const std::vector<char>& cubin, const std::string& kernel, unsigned* regs, unsigned* stack)
{
    const unsigned char* base = reinterpret_cast<const unsigned char*>(cubin.data());
    const size_t len = cubin.size();
    if(len < sizeof(Elf64Ehdr) || memcmp(base, "\177ELF", 4) != 0 || base[4] != 2)
        return false;
    Elf64Ehdr eh;
    memcpy(&eh, base, sizeof(eh));
    if(eh.shoff == 0 || eh.shentsize != sizeof(Elf64Shdr) || eh.shoff + (uint64_t)eh.shnum*sizeof(Elf64Shdr) > len
       || eh.shstrndx >= eh.shnum)
        return false;
    std::vector<Elf64Shdr> sh(eh.shnum);
    memcpy(sh.data(), base + eh.shoff, eh.shnum*sizeof(Elf64Shdr));
    // symbol index of the kernel
    long symidx = -1;
    for(const Elf64Shdr& st : sh)
    {
        if(st.type != 2 /* SHT_SYMTAB */ || st.link >= sh.size() || st.entsize != sizeof(Elf64Sym))
            continue;
        const Elf64Shdr& str = sh[st.link];
        if(st.offset + st.size > len || str.offset + str.size > len)
            return false;
        for(size_t i = 0; i < st.size/sizeof(Elf64Sym); ++i)
        {
            Elf64Sym sym;
            memcpy(&sym, base + st.offset + i*sizeof(Elf64Sym), sizeof(sym));
            if(sym.name < str.size && (sym.info & 0xf) == 2 /* STT_FUNC */
               && kernel == reinterpret_cast<const char*>(base + str.offset + sym.name))
                symidx = (long)i;
        }
    }
    if(symidx < 0)
        return false;
    const Elf64Shdr& names = sh[eh.shstrndx];
    bool have_regs = false, have_stack = false;
    for(const Elf64Shdr& sec : sh)
    {
        if(sec.name >= names.size || sec.offset + sec.size > len)
            continue;
        if(strcmp(reinterpret_cast<const char*>(base + names.offset + sec.name), ".nv.info") != 0)
            continue;
        const unsigned char* d = base + sec.offset;
        for(size_t i = 0; i + 4 <= sec.size; )
        {
            const unsigned fmt = d[i], attr = d[i + 1];
            if(fmt != 4)        // formats 1-3 carry their value in the 4-byte record itself
            {
                i += 4;
                continue;
            }
            uint16_t sz;
            memcpy(&sz, d + i + 2, 2);
            if(i + 4 + sz > sec.size)
                break;
            if(sz == 8 && (attr == 0x2f || attr == 0x12))
            {
                uint32_t v[2];
                memcpy(v, d + i + 4, 8);
                if((long)v[0] == symidx)
                {
                    if(attr == 0x2f) { *regs = v[1]; have_regs = true; }
                    else { *stack = v[1]; have_stack = true; }
                }
            }
            i += 4 + sz;
        }
    }
    return have_regs && have_stack;
}

} // namespace lcu

// ---------------------------------------------------------------------------
// object cache of a context
// ---------------------------------------------------------------------------
std::vector<lcu::Header> lcu_ctx::headers() const
{
    return { { "shim.cuh", shim }, { "object.cuh", object_hdr } };
}

// The reference builds with "-cl-denorms-are-zero -cl-fast-relaxed-math"
// (src/lensed.c:744-748).  Default here: IEEE division / square root, accurate
// libdevice transcendentals and no FMA contraction, so that object code rounds
// like the CPU oracle.  Opt-in relaxations (model flags): LCU_FAST_MATH = FMA
// contraction, LCU_FAST_INTRINSICS = hardware exp2/log2/sin/cos approximations
// for exp/log/pow/sin/cos in source and foreground objects (remapped at source
// level by shim.cuh; measured as accurate as the strict build on Sersic
// scenes), LCU_FAST_LENS_INTRINSICS = the same in lens objects (costs accuracy
// in the deflection), LCU_FAST_ATANH = atanh of lens objects on the hardware
// log2 (absolute error 2e-7).  Division and square root are always IEEE.
// Denormals are flushed either way.
std::vector<std::string> lcu_ctx::build_options(unsigned flags) const
{
    std::vector<std::string> o = {
        "--gpu-architecture=sm_100a",
        "--std=c++17",
        "--device-as-default-execution-space",
        "--generate-line-info",
        "-diag-suppress=177,550",
    };
    o.push_back((flags & LCU_FAST_INTRINSICS) ? "-DLCU_INTRINSICS_SOURCE=1" : "-DLCU_INTRINSICS_SOURCE=0");
    o.push_back((flags & LCU_FAST_LENS_INTRINSICS) ? "-DLCU_INTRINSICS_LENS=1" : "-DLCU_INTRINSICS_LENS=0");
    o.push_back((flags & LCU_FAST_ATANH) ? "-DLCU_ATANH_LENS=1" : "-DLCU_ATANH_LENS=0");
    o.push_back("-DLCU_ATANH_SOURCE=0");
    o.push_back("--ftz=true");
    o.push_back("--prec-div=true");
    o.push_back("--prec-sqrt=true");
    o.push_back((flags & LCU_FAST_MATH) ? "--fmad=true" : "--fmad=false");
    o.push_back((flags & LCU_FAST_MATH) ? "-DLCU_FMAD=1" : "-DLCU_FMAD=0");
    const char* extra = getenv("LCU_NVRTC_FLAGS");
    if(extra && *extra)
    {
        std::istringstream ss(extra);
        std::string tok;
        while(ss >> tok)
            o.push_back(tok);
    }
    return o;
}

bool lcu_ctx::compile(const std::string& source, unsigned flags, std::vector<char>* cubin, std::string* log)
{
    std::string key;
    for(const std::string& o : build_options(flags))
        key += o + '\n';
    key += source;
    auto it = cubins.find(key);
    if(it != cubins.end())
    {
        *cubin = it->second;
        log->clear();
        return true;
    }
    if(!lcu::compile_cubin(source, headers(), build_options(flags), cubin, log))
        return false;
    cubins.emplace(std::move(key), *cubin);
    return true;
}

const lcu::ObjectInfo* lcu_ctx::object(const std::string& name)
{
    using namespace lcu;
    auto it = objects.find(name);
    if(it != objects.end())
        return &it->second;

    if(name.empty() || name.find('/') != std::string::npos)
    {
        set_error("invalid object name \"%s\"", name.c_str());
        return nullptr;
    }

    bool ok = false;
    const std::string path = objects_dir + "/" + name + ".cl";
    const std::string text = read_text_file(path, &ok);
    if(!ok)
    {
        // same wording as the reference, src/kernel.c:757-759
        set_error("could not load object \"%s\" (file not found: %s)", name.c_str(), path.c_str());
        return nullptr;
    }

    ObjectInfo info;
    info.name = name;
    info.ident = make_ident(name);
    info.wrapped = wrap_object(name, info.ident, text);

    // the reference builds object_program() with all image options zero
    // (src/input/objects.c:89-91); so do we
    std::string src;
    src += "#define IMAGE_SIZE 0\n#define IMAGE_WIDTH 0\n#define IMAGE_HEIGHT 0\n"
           "#define PSF 0\n#define PSF_WIDTH 0\n#define PSF_HEIGHT 0\n#define QUAD_POINTS 0\n";
    src += "#include \"shim.cuh\"\n#include \"object.cuh\"\n";
    {
        std::string w = info.wrapped;
        size_t pos;
        while((pos = w.find("@KIND@")) != std::string::npos)
            w.replace(pos, 6, "SOURCE");
        src += w;
    }

    // first with the two-rays-per-thread copy of the per-ray functions; an
    // object whose text cannot be typed as pairs is not an error, it renders
    // one ray per thread
    std::vector<char> cubin;
    std::string log;
    info.pairable = compile_cubin("#define LCU_PAIR 1\n" + src, headers(), build_options(0), &cubin, &log);
    if(!info.pairable)
    {
        info.pair_log = log;
        if(!compile_cubin("#define LCU_PAIR 0\n" + src, headers(), build_options(0), &cubin, &log))
        {
            set_error("object %s: failed to build program\n%s", name.c_str(), log.c_str());
            return nullptr;
        }
    }

    const unsigned char* bytes = nullptr;
    size_t size = 0;
    if(!cubin_symbol(cubin, "lcu_meta_" + info.ident, &bytes, &size) || size != 3*sizeof(uint32_t))
    {
        set_error("object %s: compiled module has no metadata", name.c_str());
        return nullptr;
    }
    uint32_t meta[3];
    memcpy(meta, bytes, sizeof(meta));
    info.type = (int)meta[0];
    info.bytes = meta[1];
    // size in 4-byte words, rounding up: src/input/objects.c:139
    info.words = info.bytes/4 + (info.bytes%4 ? 1 : 0);
    {
        // which relaxed-math switch governs this object: lenses have their own
        const std::string kind = info.type == LCU_LENS ? "LENS" : "SOURCE";
        size_t pos;
        while((pos = info.wrapped.find("@KIND@")) != std::string::npos)
            info.wrapped.replace(pos, 6, kind);
    }
    if(info.type != LCU_LENS && info.type != LCU_SOURCE && info.type != LCU_FOREGROUND)
    {
        // src/input/objects.c:147-148
        set_error("object %s: invalid type (should be LENS, SOURCE or FOREGROUND)", name.c_str());
        return nullptr;
    }
    const size_t npar = meta[2];
    if(npar > 0)
    {
        if(!cubin_symbol(cubin, "lcu_parlst_" + info.ident, &bytes, &size) || size != npar*sizeof(lcu_param))
        {
            set_error("object %s: compiled module has no parameter list", name.c_str());
            return nullptr;
        }
        info.params.resize(npar);
        memcpy(info.params.data(), bytes, size);
        for(lcu_param& p : info.params)
            p.name[15] = '\0';
    }

    // the entry point the type calls for must exist (checked by a second,
    // cheap look at the text: the compile above would not notice)
    const char* fn = info.type == LCU_LENS ? "deflection" : info.type == LCU_SOURCE ? "brightness" : "foreground";
    if(text.find(fn) == std::string::npos || text.find("set") == std::string::npos)
    {
        set_error("object %s: missing %s() or set() function", name.c_str(), fn);
        return nullptr;
    }

    auto res = objects.emplace(name, std::move(info));
    return &res.first->second;
}
