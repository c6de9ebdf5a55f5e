/*
 * host_check.c -- a C host for the C ABI, written the way the reference's own
 * host code would use it: metadata as in add_object()
 * (src/input/objects.c:72-239), device set-up as in src/lensed.c:644-1112, one
 * evaluation as in loglike() (src/nested.c:63-115).  Built and run by
 * tests/test_c_host.py.
 *
 *   host_check meta                      object metadata + quadrature rules
 *   host_check loglike <device> <size>   a lens + source model on a blank image
 *   host_check latency <device> <size> <n>   n one-point calls in a row, as MultiNest would make them
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lensed_cuda.h"

#define CHECK(call) do { if((call) != LCU_OK) { fprintf(stderr, "%s: %s\n", #call, lcu_last_error()); return 1; } } while(0)

static const char* OBJECTS[] = { "sis", "sis_plus_shear", "sie", "sie_plus_shear", "nsis", "nsie", "point_mass", "epl",
                                 "epl_plus_shear", "sersic", "sersic-old", "devauc", "exponential", "gauss", "sky" };

static int meta(void)
{
    lcu_ctx* ctx;
    CHECK(lcu_create(-1, NULL, NULL, &ctx));
    for(size_t i = 0; i < sizeof(OBJECTS)/sizeof(OBJECTS[0]); ++i)
    {
        int type; size_t words, npar;
        lcu_param pars[16];
        CHECK(lcu_object_info(ctx, OBJECTS[i], &type, &words, &npar, pars, 16));
        const char* why = NULL;
        const int pair = lcu_object_pairable(ctx, OBJECTS[i], &why);
        if(pair != 1)
        {
            fprintf(stderr, "%s does not compile for two rays per thread: %s\n", OBJECTS[i], why ? why : "?");
            return 1;
        }
        printf("object %s %c %zu %zu", OBJECTS[i], type, words, npar);
        for(size_t j = 0; j < npar; ++j)
            printf(" %s:%d:%g:%g:%d", pars[j].name, pars[j].type, pars[j].bounds[0], pars[j].bounds[1],
                   pars[j].defval > 0 || signbit(pars[j].defval));
        printf("\n");
    }
    for(int r = 0; r < lcu_quad_rule_count(); ++r)
        printf("rule %s %d\n", lcu_quad_rule_name(r), lcu_quad_rule(lcu_quad_rule_name(r), 1, 1, NULL, NULL));
    /* an unknown object is an error code + message, never exit() */
    {
        int type; size_t words, npar;
        if(lcu_object_info(ctx, "nonesuch", &type, &words, &npar, NULL, 0) != LCU_E_IO)
            return 1;
        printf("error %s\n", lcu_last_error());
    }
    lcu_destroy(ctx);
    return 0;
}

static int loglike(int device, size_t size)
{
    lcu_ctx* ctx;
    lcu_model* model;
    CHECK(lcu_create(device, NULL, NULL, &ctx));

    /* [objects] lens = sie, source = sersic; source position with "image" priors */
    int ipp_src[7] = { 1, 1, 0, 0, 0, 0, 0 };
    lcu_object_spec objs[2] = { { "sie", NULL }, { "sersic", ipp_src } };

    int nq = lcu_quad_rule("g3k7", 1, 1, NULL, NULL);
    float* qq = malloc(2*nq*sizeof(float));
    float* ww = malloc(2*nq*sizeof(float));
    lcu_quad_rule("g3k7", 1, 1, qq, ww);

    float* image = calloc(size*size, sizeof(float));
    float* weight = malloc(size*size*sizeof(float));
    for(size_t i = 0; i < size*size; ++i)
        weight[i] = 1.0f + (float)(i % 7);
    float psf[9] = { 0.05f, 0.1f, 0.05f, 0.1f, 0.4f, 0.1f, 0.05f, 0.1f, 0.05f };

    lcu_model_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.width = desc.height = size;
    desc.pcs[0] = desc.pcs[1] = desc.pcs[2] = desc.pcs[3] = 1;
    desc.nq = (size_t)nq; desc.qq = qq; desc.ww = ww;
    desc.image = image; desc.weight = weight;
    desc.psf = psf; desc.psf_width = desc.psf_height = 3;
    CHECK(lcu_model_create(ctx, objs, 2, &desc, &model));
    printf("model npars %zu words %zu\n", lcu_model_npars(model), lcu_model_words(model));

    const float c = 0.5f*(float)(size + 1);
    float params[3][12];
    for(int b = 0; b < 3; ++b)
    {
        float p[12] = { c, c, 0.2f*(float)size, 0.75f + 0.05f*(float)b, 45.f,           /* lens x y r q pa */
                        c + 0.25f*(float)size, c + 0.1f*(float)size, 0.04f*(float)size, -3.f, 2.f, 0.8f, 30.f };
        memcpy(params[b], p, sizeof(p));
    }

    double one, three[3];
    CHECK(lcu_loglike(model, params[0], &one));
    CHECK(lcu_loglike_batch(model, 3, &params[0][0], three));
    printf("lnew %.17g %.17g %.17g %.17g\n", one, three[0], three[1], three[2]);

    float* img = malloc(size*size*sizeof(float));
    CHECK(lcu_render(model, params[0], img, NULL, NULL, NULL));
    double sum = 0;
    for(size_t i = 0; i < size*size; ++i)
        sum += img[i];
    printf("flux %.9g launches %llu\n", sum, lcu_launch_count());

    /* data preparation on the device: weight = gain/(image + offset), one masked pixel;
       then the same evaluation against the new map */
    {
        int* mask = calloc(size*size, sizeof(int));
        float* w2 = malloc(size*size*sizeof(float));
        double again;
        mask[size + 1] = 1;
        CHECK(lcu_model_make_weight(model, NULL, 1800.f, 2.9633, mask));
        CHECK(lcu_model_get_weight(model, w2));
        CHECK(lcu_loglike(model, params[0], &again));
        printf("prep %d %.9g %.9g %.17g\n", lcu_model_rays_per_thread(model), w2[0], w2[size + 1], again);
        CHECK(lcu_model_set_data(model, NULL, weight));
        CHECK(lcu_loglike(model, params[0], &again));
        printf("back %.17g\n", again);
        free(mask); free(w2);
    }

    lcu_model_destroy(model);
    lcu_destroy(ctx);
    free(qq); free(ww); free(image); free(weight); free(img);
    return 0;
}

/* the sampler's callback pattern (src/nested.c:63-115): one point per call, the
   next call needs the previous result; prints the mean time per call */
static int latency(int device, size_t size, int n)
{
    lcu_ctx* ctx;
    lcu_model* model;
    CHECK(lcu_create(device, NULL, NULL, &ctx));
    lcu_object_spec objs[2] = { { "sie", NULL }, { "sersic", NULL } };
    int nq = lcu_quad_rule("g3k7", 1, 1, NULL, NULL);
    float* qq = malloc(2*nq*sizeof(float));
    float* ww = malloc(2*nq*sizeof(float));
    lcu_quad_rule("g3k7", 1, 1, qq, ww);
    float* image = calloc(size*size, sizeof(float));
    float* weight = malloc(size*size*sizeof(float));
    for(size_t i = 0; i < size*size; ++i)
        weight[i] = 1.0f;
    float psf[81];
    for(int i = 0; i < 81; ++i)
        psf[i] = 1.0f/81;
    lcu_model_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.width = desc.height = size;
    desc.pcs[0] = desc.pcs[1] = desc.pcs[2] = desc.pcs[3] = 1;
    desc.nq = (size_t)nq; desc.qq = qq; desc.ww = ww;
    desc.image = image; desc.weight = weight;
    desc.psf = psf; desc.psf_width = desc.psf_height = 9;
    desc.flags = LCU_FAST_INTRINSICS | LCU_FAST_ATANH;
    CHECK(lcu_model_create(ctx, objs, 2, &desc, &model));
    const float c = 0.5f*(float)(size + 1);
    float p[12] = { c, c, 0.2f*(float)size, 0.75f, 45.f, c + 2.f, c + 1.f, 0.04f*(float)size, -3.f, 2.f, 0.8f, 30.f };
    double lnew = 0, sum = 0;
    for(int i = 0; i < 20; ++i)
        CHECK(lcu_loglike(model, p, &lnew));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for(int i = 0; i < n; ++i)
    {
        p[3] = 0.75f + 1e-4f*(float)(i & 15);
        CHECK(lcu_loglike(model, p, &lnew));
        sum += lnew;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double us = ((double)(t1.tv_sec - t0.tv_sec)*1e9 + (double)(t1.tv_nsec - t0.tv_nsec))/1e3/n;
    printf("latency %.3f us per lcu_loglike (%d calls, %zux%zu, sie + sersic, 9x9 PSF, g3k7) checksum %.9g\n", us, n, size, size, sum);
    /* the same points with two evaluations in flight (lcu_loglike_async / _wait): what a host
       that prepares point i + 1 while point i is evaluated sees per point */
    double sum2 = 0;
    int ticket = -1, next;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for(int i = 0; i < n; ++i)
    {
        p[3] = 0.75f + 1e-4f*(float)(i & 15);
        CHECK(lcu_loglike_async(model, p, &next));
        if(ticket >= 0)
        {
            CHECK(lcu_loglike_wait(model, ticket, &lnew));
            sum2 += lnew;
        }
        ticket = next;
    }
    CHECK(lcu_loglike_wait(model, ticket, &lnew));
    sum2 += lnew;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double us2 = ((double)(t1.tv_sec - t0.tv_sec)*1e9 + (double)(t1.tv_nsec - t0.tv_nsec))/1e3/n;
    printf("pipelined %.3f us per point with two in flight (lcu_loglike_async / lcu_loglike_wait) checksum %.9g %s\n",
           us2, sum2, sum2 == sum ? "same" : "DIFFERENT");
    lcu_model_destroy(model);
    lcu_destroy(ctx);
    free(qq); free(ww); free(image); free(weight);
    return 0;
}

int main(int argc, char* argv[])
{
    if(argc >= 2 && strcmp(argv[1], "meta") == 0)
        return meta();
    if(argc >= 4 && strcmp(argv[1], "loglike") == 0)
        return loglike(atoi(argv[2]), (size_t)atol(argv[3]));
    if(argc >= 5 && strcmp(argv[1], "latency") == 0)
        return latency(atoi(argv[2]), (size_t)atol(argv[3]), atoi(argv[4]));
    fprintf(stderr, "usage: host_check meta | loglike <device> <size> | latency <device> <size> <n>\n");
    return 2;
}
