#!/usr/bin/env python
"""Build oracle/_ref/: the reference's own implementation of the hot path,
compiled on the host from the sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing is copied into the repository: generated
intermediate text lives in a temporary directory, only shared libraries are
written to oracle/_ref/ (git-ignored, travels to the GPU box).

Pipeline
  1. gcc: reference src/kernel.c + src/log.c + oracle/refgen.c -> refgen, a
     tool that prints the program text the reference's kernel assembler
     produces (object_program / main_program, src/kernel.c:818-879).
  2. for every object in <reference>/objects: refgen object <name> -> text ->
     vector-literal rewrite -> g++ with oracle/ref_shim.h -> run the reference's
     meta_<name> / params_<name> kernels on the host -> type, size, parameters.
  3. for every model configuration in oracle/ref_configs.json: refgen main ... ->
     text (reference object.cl, constants.cl, objects, generated compute() and
     set_params(), lensed.cl) -> rewrite -> one namespace per configuration.
  4. g++: all of the above + reference src/quadrature.c and src/quad/ tables +
     oracle/ref_driver.cpp -> oracle/_ref/liblensed_ref.so (strict float32:
     -O2 -ffp-contract=off) and liblensed_ref_fast.so (-O3 -ffast-math, AVX2,
     OpenMP; the timed CPU baseline, kind "reference").

  5. the same plugin text a second way, for timing only: per-ray functions and
     the generated compute() with float = 8 consecutive work-items (AVX2 lanes,
     oracle/ref_shim_simd.h), uniform data kept scalar, render / convolve /
     loglike driven over work-groups > 1 (ref_driver.cpp -DREF_SIMD) ->
     liblensed_ref_simd.so (AVX2, 8 lanes) and liblensed_ref_simd512.so (AVX-512,
     16 lanes; used where the CPU has it): what a vectorising OpenCL CPU runtime (the Intel
     runtime of .travis.yml:60-68) makes of the reference's kernels.

The only transformation applied to reference text is the token rewrite of
OpenCL vector literals "(float2)(a, b)" -> "float2(a, b)" (a cast of a comma
expression in C++), and -fpermissive for the implicit void* conversions of the
generated "(local void*)(data + N)" arguments.
"""
import ctypes as C
import glob
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LENSED_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CC = os.environ.get("ORACLE_CC", "gcc")
CXX = os.environ.get("ORACLE_CXX", "g++")

LITERAL = re.compile(r"\(\s*(float2|float4|mat22)\s*\)\s*\(")


def run(cmd, **kw):
    print("+", " ".join(cmd[:6]), "..." if len(cmd) > 6 else "", flush=True)
    return subprocess.run(cmd, check=True, **kw)


def rewrite(text):
    return LITERAL.sub(r"\1(", text)


def ident(name):
    return re.sub(r"[^A-Za-z0-9_]", "_", name)


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", lambda m: re.sub(r"[^\n]", " ", m.group(0)), text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def match_brace(text, i):
    """index just past the brace that closes the one at text[i]"""
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced braces")


def object_parts(name):
    """(data block, per-ray text) of an object file: `data { ... };` as written,
    and the file with type / params / data / set() blanked out -- what is left is
    the deflection / brightness / foreground function (and any helpers)."""
    text = strip_comments(open(os.path.join(REF, "objects", name + ".cl")).read())
    out = text
    data_block = None
    for key in ("type", "params", "data"):
        m = re.search(r"(?m)^\s*%s\b" % key, out)
        if not m:
            continue
        if key == "type":
            end = out.index(";", m.start()) + 1
        else:
            end = match_brace(out, out.index("{", m.start()))
            end = out.index(";", end) + 1
        if key == "data":
            data_block = out[m.start():end]
        out = out[:m.start()] + out[end:]
    m = re.search(r"static\s+void\s+set\s*\(", out)
    if m:
        end = match_brace(out, out.index("{", m.start()))
        out = out[:m.start()] + out[end:]
    if data_block is None:
        raise ValueError(f"{name}: no data block")
    return data_block, out


def compute_text(main_text):
    """the generated compute() of a main program (src/kernel.c:65-111)"""
    i = main_text.index("static float compute(")
    return main_text[i:match_brace(main_text, main_text.index("{", i))]


SELECT = re.compile(r"dot\(a,\s*a\) < HUGE_VALF \? a : float2\(1E10f, 1E10f\)")


def simd_unit(ci, cfg, main_text):
    """Translation unit of one configuration for the SIMD build: scalar data
    structs, then the per-ray functions and compute() with float = vfloat, then
    the render loop of kernel/lensed.cl:9-38 over 8 work-items at a time."""
    kdir = os.path.join(REF, "kernel")
    names = list(dict.fromkeys(cfg["objects"]))
    o = ['#include "ref_shim.h"\n#include "ref_shim_simd.h"\n', f"namespace simd_cfg_{ci} {{\n"]
    o.append(rewrite(open(os.path.join(kdir, "object.cl")).read()))
    o.append(rewrite(open(os.path.join(kdir, "constants.cl")).read()))
    parts = {n: object_parts(n) for n in names}
    for n in names:
        o.append(f"#define data struct data_{ident(n)}\n{rewrite(parts[n][0])}\n#undef data\n")
    o.append("#define float vfloat\n#define float2 vfloat2\n#define float4 vfloat4\n#define mat22 vfloat4\n")
    for n in names:
        o.append(f"#define data struct data_{ident(n)}\n")
        for fn in ("deflection", "brightness", "foreground"):
            o.append(f"#define {fn} {fn}_{ident(n)}\n")
        o.append(rewrite(parts[n][1]) + "\n")
        o.append("#undef data\n#undef deflection\n#undef brightness\n#undef foreground\n")
    comp, nsub = SELECT.subn("ref_select(dot(a,a) < HUGE_VALF, a, float2(1E10f, 1E10f))", rewrite(compute_text(main_text)))
    if "deflection_" in comp and nsub == 0:
        raise ValueError("generated compute(): deflection guard not found")
    o.append(comp + "\n")
    o.append("#undef float\n#undef float2\n#undef float4\n#undef mat22\n")
    o.append("""
static void render8(const uint* data, const float* pcs, const ::float2* qq, const ::float2* ww, int nq,
                    long k0, long k1, long size, int width, float* value, float* error)
{
    for(long k = k0; k < k1; k += REF_LANES)
    {
        vfloat px, py;
        for(int l = 0; l < REF_LANES; ++l)
        {
            const long kk = k + l < size ? k + l : size - 1;
            px.v[l] = (float)(kk % width);
            py.v[l] = (float)(kk / width);
        }
        const vfloat2 x(vfloat(pcs[0]) + vfloat(pcs[2])*px, vfloat(pcs[1]) + vfloat(pcs[3])*py);
        vfloat f0 = 0.0f, f1 = 0.0f;
        for(int n = 0; n < nq; ++n)
        {
            const vfloat c = compute((uint*)data, x + qq[n]);
            f0 += vfloat(ww[n].x)*c;
            f1 += vfloat(ww[n].y)*c;
        }
        for(int l = 0; l < REF_LANES && k + l < k1; ++l)
        {
            value[k + l] = f0.v[l];
            error[k + l] = f1.v[l];
        }
    }
}
""")
    o.append("}\n#undef kernel\n#undef global\n#undef local\n#undef constant\n#undef this\n")
    o.append(f"extern const ref_simd_render_fn REF_SIMD_RENDER_{ci} = simd_cfg_{ci}::render8;\n")
    return "".join(o)


def main():
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} not present, nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="lensed_ref_")
    env = dict(os.environ, LENSED_PATH=REF.rstrip("/") + "/")
    src = os.path.join(REF, "src")
    try:
        # 1. the reference's kernel assembler as a command-line tool
        refgen = os.path.join(tmp, "refgen")
        # -Dsnprintf=ref_snprintf: see the comment in refgen.c (glibc clips the
        # reference's "unlimited" snprintf calls by one character)
        run([CC, "-std=c99", "-D_GNU_SOURCE", "-O1", "-fno-builtin", "-Dsnprintf=ref_snprintf", "-w", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-I", src,
             os.path.join(src, "kernel.c"), os.path.join(src, "log.c"), os.path.join(HERE, "refgen.c"),
             "-lm", "-o", refgen])

        def gen(*args):
            return subprocess.run([refgen, *args], check=True, env=env, capture_output=True, text=True).stdout

        # 2. object metadata through the reference's own meta kernels
        names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(REF, "objects", "*.cl")))
        # the reference pastes the object name into identifiers
        # (src/kernel.c:153-162): names that are not identifiers cannot be built by it
        usable = [n for n in names if ident(n) == n]
        skipped = [n for n in names if n not in usable]
        meta_units = []
        for n in usable:
            path = os.path.join(tmp, f"meta_{n}.cpp")
            with open(path, "w") as f:
                f.write('#include "ref_shim.h"\n')
                f.write(f"namespace meta_ns_{n} {{\n{rewrite(gen('object', n))}\n}}\n")
            meta_units.append(path)
        table = os.path.join(tmp, "objects_table.cpp")
        with open(table, "w") as f:
            f.write('#include "ref_shim.h"\n#undef kernel\n#undef global\n#undef local\n#undef constant\n#undef this\n')
            for n in usable:
                f.write(f"namespace meta_ns_{n} {{ void meta_{n}(int*, ulong*, ulong*); "
                        f"void params_{n}(char16*, int*, float2*, float*); }}\n")
            f.write("extern const ref_object REF_OBJECTS[] = {\n")
            for n in usable:
                f.write(f'    {{ "{n}", meta_ns_{n}::meta_{n}, meta_ns_{n}::params_{n} }},\n')
            f.write(f"}};\nextern const int REF_NOBJECTS = {len(usable)};\n")
            f.write("extern const int REF_NCONFIGS_STAGE = 0;\n")
        strict = ["-std=gnu++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fpermissive", "-w", "-fPIC", "-fopenmp",
                  "-I", HERE]
        stage = os.path.join(tmp, "libstage.so")
        stub = os.path.join(tmp, "stage_stub.cpp")
        with open(stub, "w") as f:
            f.write('#include "ref_shim.h"\n#undef kernel\n#undef global\n#undef local\n#undef constant\n#undef this\n'
                    "thread_local ref_workitem ref_wi;\n"
                    "int ref_image_size, ref_image_width, ref_image_height, ref_psf, ref_psf_width, ref_psf_height, ref_quad_points;\n"
                    "extern const ref_object REF_OBJECTS[]; extern const int REF_NOBJECTS;\n"
                    'extern "C" int stage_meta(int i, char* name, int* type, unsigned long* size, unsigned long* npar, int* ptypes) {\n'
                    "  if(i >= REF_NOBJECTS) return 1; strcpy(name, REF_OBJECTS[i].name);\n"
                    "  REF_OBJECTS[i].meta(type, size, npar);\n"
                    "  char16 nm[16]; float2 b[16]; float d[16];\n"
                    "  for(unsigned long j = 0; j < *npar; ++j) { ref_wi.gid[0] = j; REF_OBJECTS[i].params(nm, ptypes, b, d); }\n"
                    "  return 0; }\n")
        run([CXX, *strict, "-shared", "-o", stage, stub, table, *meta_units])
        L = C.CDLL(stage)
        metas = {}
        for i in range(len(usable)):
            nm = C.create_string_buffer(64)
            t, s, k = C.c_int(), C.c_ulong(), C.c_ulong()
            pt = (C.c_int*16)()
            assert L.stage_meta(i, nm, C.byref(t), C.byref(s), C.byref(k), pt) == 0
            words = s.value//4 + (1 if s.value % 4 else 0)      # src/input/objects.c:139
            metas[nm.value.decode()] = dict(type=chr(t.value), words=words, npars=k.value,
                                            ptypes="".join(str(pt[j]) for j in range(k.value)))
        print("objects:", {k: (v["type"], v["words"], v["npars"]) for k, v in metas.items()}, "skipped:", skipped)

        # 3. main programs
        configs = json.load(open(os.path.join(HERE, "ref_configs.json")))
        cfg_units = []
        simd_units = []
        for ci, cfg in enumerate(configs):
            specs = []
            for name, ipp in zip(cfg["objects"], cfg["ipp"]):
                m = metas[name]
                ipp = (ipp or "").ljust(m["npars"], "0") if m["npars"] else "0"
                specs.append(f"{name}:{m['type']}:{m['words']}:{m['npars']}:{ipp}:{m['ptypes'] or '0'}")
            main_raw = gen("main", *specs)
            text = rewrite(main_raw)
            spath = os.path.join(tmp, f"simd_{ci}.cpp")
            with open(spath, "w") as f:
                f.write(simd_unit(ci, cfg, main_raw))
            simd_units.append(spath)
            path = os.path.join(tmp, f"cfg_{ci}.cpp")
            with open(path, "w") as f:
                f.write('#include "ref_shim.h"\n')
                f.write(f"namespace cfg_{ci} {{\n{text}\n}}\n")
                f.write("#undef kernel\n#undef global\n#undef local\n#undef constant\n#undef this\n")
                f.write(f"extern const ref_program REF_PROGRAM_{ci} = {{ cfg_{ci}::set_params, cfg_{ci}::render, "
                        f"cfg_{ci}::loglike, cfg_{ci}::convolve }};\n")
            cfg_units.append(path)
        ctable = os.path.join(tmp, "configs_table.cpp")
        with open(ctable, "w") as f:
            f.write('#include "ref_shim.h"\n#undef kernel\n#undef global\n#undef local\n#undef constant\n#undef this\n')
            f.write("struct ref_config { int nobjs; const char* names[8]; const char* ipp[8]; const ref_program* program; };\n")
            for ci in range(len(configs)):
                f.write(f"extern const ref_program REF_PROGRAM_{ci};\n")
            f.write("extern const ref_config REF_CONFIGS[] = {\n")
            for ci, cfg in enumerate(configs):
                nm = ", ".join(f'"{n}"' for n in cfg["objects"])
                ip = ", ".join('"%s"' % (i or "").ljust(metas[n]["npars"], "0") for n, i in zip(cfg["objects"], cfg["ipp"]))
                f.write(f"    {{ {len(cfg['objects'])}, {{ {nm} }}, {{ {ip} }}, &REF_PROGRAM_{ci} }},\n")
            f.write(f"}};\nextern const int REF_NCONFIGS = {len(configs)};\n")

        # 4. link, twice.  The quadrature rules are the reference's C files,
        #    compiled in place against a one-typedef stand-in for <CL/cl.h>.
        stubinc = os.path.join(tmp, "stubinc", "CL")
        os.makedirs(stubinc)
        with open(os.path.join(stubinc, "cl.h"), "w") as f:
            f.write("#pragma once\n#include <stddef.h>\ntypedef struct { float s[2]; } cl_float2;\n"
                    "typedef void* cl_platform_id; typedef void* cl_device_id; typedef void* cl_context; typedef unsigned long cl_device_type;\n")
        quad_objs = []
        for cfile in [os.path.join(src, "quadrature.c")] + sorted(glob.glob(os.path.join(src, "quad", "*.c"))):
            o = os.path.join(tmp, "q_" + os.path.basename(cfile) + ".o")
            run([CC, "-std=c99", "-O2", "-fPIC", "-w", "-I", os.path.join(tmp, "stubinc"), "-I", src, "-c", cfile, "-o", o])
            quad_objs.append(o)
        units = [os.path.join(HERE, "ref_driver.cpp"), table, ctable, *meta_units, *cfg_units]
        from concurrent.futures import ThreadPoolExecutor

        def compile_link(name, flags, srcs, libs=()):
            """one object file per unit, compiled in parallel, then linked into oracle/_ref/<name>"""
            tag = os.path.splitext(name)[0]
            objs = [os.path.join(tmp, f"{tag}_{i}.o") for i in range(len(srcs))]
            print(f"+ {CXX} {' '.join(flags[:5])} ... -> {name} ({len(srcs)} units, {os.cpu_count()} jobs)", flush=True)
            with ThreadPoolExecutor(os.cpu_count() or 1) as pool:
                list(pool.map(lambda so: subprocess.run([CXX, *flags, "-c", so[0], "-o", so[1]], check=True), zip(srcs, objs)))
            run([CXX, "-shared", "-fopenmp", "-o", os.path.join(OUT, name), *objs, *quad_objs, *libs, "-lm"])
        compile_link("liblensed_ref.so", strict, units)
        fast = ["-std=gnu++17", "-O3", "-march=x86-64-v3", "-ffast-math", "-fpermissive", "-w", "-fPIC", "-fopenmp", "-I", HERE]
        compile_link("liblensed_ref_fast.so", fast, units)
        # 5. the vectorised build (timing only)
        stable = os.path.join(tmp, "simd_table.cpp")
        with open(stable, "w") as f:
            f.write('#include "ref_shim.h"\n#include "ref_shim_simd.h"\n#undef kernel\n#undef global\n#undef local\n#undef constant\n#undef this\n')
            for ci in range(len(configs)):
                f.write(f"extern const ref_simd_render_fn REF_SIMD_RENDER_{ci};\n")
            f.write("extern const ref_simd_render_fn REF_SIMD_RENDER[] = {\n")
            f.write("".join(f"    REF_SIMD_RENDER_{ci},\n" for ci in range(len(configs))))
            f.write("};\n")
        for tag, lanes, march in (("simd", 8, "x86-64-v3"), ("simd512", 16, "x86-64-v4")):
            flags = ["-std=gnu++17", "-O3", f"-march={march}", "-ffast-math", "-fpermissive", "-w", "-fPIC", "-fopenmp",
                     "-I", HERE, f"-DREF_SIMD={lanes}"]
            compile_link(f"liblensed_ref_{tag}.so", flags, [*units, stable, *simd_units], libs=["-lmvec"])
        json.dump(dict(objects=metas, skipped=skipped, configs=configs), open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
        print("built", os.listdir(OUT))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
