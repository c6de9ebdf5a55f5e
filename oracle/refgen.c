/*
 * refgen.c -- drives the REFERENCE's own kernel assembler (src/kernel.c,
 * compiled from where it lies, together with src/log.c) to emit the program
 * text Lensed would hand to the OpenCL compiler.  TEST INFRASTRUCTURE ONLY;
 * built and run by oracle/build_ref.py, never shipped.
 *
 *   refgen object <name>
 *       object_program(): object.cl, constants.cl, the object file in its
 *       name-mangling macros, meta_<name> and params_<name> kernels
 *   refgen main <name>:<type>:<words>:<npars>:<ipp-bits>:<partypes> ...
 *       main_program() for the object list in ini order; <ipp-bits> and
 *       <partypes> are one character per parameter ('0'/'1', and the
 *       parameter type digit of src/input.h:12-21)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "input.h"
#include "kernel.h"

/*
 * src/kernel.c:455-470 formats its second pass with snprintf(out, (size_t)-1,
 * ...), i.e. "no limit".  glibc >= 2.37 clips such a call one character short
 * (buffer end wraps around the address space), which splits the generated
 * set_params text with NUL bytes.  build_ref.py therefore compiles kernel.c
 * with -Dsnprintf=ref_snprintf; this wrapper gives the call the meaning the
 * reference intends.  The reference source itself is not touched.
 */
#include <stdarg.h>
#include <stdint.h>
int ref_snprintf(char* s, size_t n, const char* fmt, ...)
{
    va_list ap;
    int r;
    va_start(ap, fmt);
    if(n == (size_t)-1)
        r = vsprintf(s, fmt, ap);
    else
        r = (vsnprintf)(s, n, fmt, ap);
    va_end(ap);
    return r;
}

/* src/path.h:3 -- set from the environment instead of the executable path */
const char* LENSED_PATH = NULL;

int main(int argc, char* argv[])
{
    size_t nkernels = 0;
    const char** kernels = NULL;
    const char* root = getenv("LENSED_PATH");

    if(!root || argc < 3)
    {
        fprintf(stderr, "usage: LENSED_PATH=<root>/ refgen object <name> | main <spec>...\n");
        return 2;
    }
    LENSED_PATH = root;

    if(strcmp(argv[1], "object") == 0)
    {
        object_program(argv[2], &nkernels, &kernels);
    }
    else if(strcmp(argv[1], "main") == 0)
    {
        size_t nobjs = (size_t)(argc - 2);
        object* objs = calloc(nobjs, sizeof(object));
        for(size_t i = 0; i < nobjs; ++i)
        {
            char* spec = strdup(argv[2 + i]);
            char* name = strtok(spec, ":");
            char* type = strtok(NULL, ":");
            char* words = strtok(NULL, ":");
            char* npars = strtok(NULL, ":");
            char* ipp = strtok(NULL, ":");
            char* ptypes = strtok(NULL, ":");
            if(!name || !type || !words || !npars || !ipp || !ptypes)
            {
                fprintf(stderr, "bad object spec: %s\n", argv[2 + i]);
                return 2;
            }
            objs[i].name = name;
            objs[i].id = name;
            objs[i].type = type[0];
            objs[i].size = (size_t)atol(words);
            objs[i].npars = (size_t)atol(npars);
            objs[i].pars = calloc(objs[i].npars ? objs[i].npars : 1, sizeof(param));
            for(size_t j = 0; j < objs[i].npars; ++j)
            {
                objs[i].pars[j].ipp = ipp[j] == '1';
                objs[i].pars[j].type = ptypes[j] - '0';
            }
        }
        main_program(nobjs, objs, &nkernels, &kernels);
    }
    else
        return 2;

    for(size_t i = 0; i < nkernels; ++i)
        fputs(kernels[i], stdout);
    return 0;
}
