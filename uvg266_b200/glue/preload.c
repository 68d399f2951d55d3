/*
 * preload.c - drop the "cuda" strategies into an UNMODIFIED uvg266 binary/library.
 *
 * The reference's group registrars (uvg_strategy_register_picture / _dct / _quant,
 * strategies-picture.c:92, strategies-dct.c:75, strategies-quant.c:49) are ordinary exported
 * functions that uvg_strategyselector_init (strategyselector.c:55) calls through the PLT, so an
 * LD_PRELOADed definition runs instead; it forwards to the original and then appends the cuda
 * registrations - the same effect as the 3-line source patch shown in INTEGRATION.md.
 *
 *   LD_PRELOAD=libuvg266_b200_glue8.so uvg266 -i in.yuv ...
 *   UVG266_B200_GROUPS=picture,dct,quant[,intra][,ipol]  (default picture,dct,quant)   UVG266_B200_DISABLE=1 (off)
 *   UVG266_B200_GROUPS=hook : no leaf strategies at all, only the batched call site of a reference built with
 *                             glue/patches/search_intra_rough.patch (one launch per rough-searched CU)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int uvg_strategy_register_picture_cuda(void *opaque, uint8_t bitdepth);
int uvg_strategy_register_dct_cuda(void *opaque, uint8_t bitdepth);
int uvg_strategy_register_quant_cuda(void *opaque, uint8_t bitdepth);
int uvg_strategy_register_intra_cuda(void *opaque, uint8_t bitdepth);
int uvg_strategy_register_ipol_cuda(void *opaque, uint8_t bitdepth);
int uvg_b200_install_rough_hook_only(uint8_t bitdepth);

typedef int (*registrar_t)(void *, uint8_t);
unsigned long long uvg_b200_launch_count(void);

/* UVG266_B200_VERBOSE=1: report at exit how many kernels the encode launched (the tests use it
 * to prove the cuda strategies, not the CPU ones, produced the bitstream). */
__attribute__((destructor)) static void report(void)
{
  if (getenv("UVG266_B200_VERBOSE"))
    fprintf(stderr, "uvg266_b200: kernel launches = %llu\n", uvg_b200_launch_count());
}

static int group_enabled(const char *name)
{
  if (getenv("UVG266_B200_DISABLE")) return 0;
  const char *g = getenv("UVG266_B200_GROUPS");
  if (!g) g = "picture,dct,quant";                    /* intra and ipol are opt-in: one launch per predicted block */
  return strstr(g, name) != NULL;
}

static int forward(const char *sym, const char *group, registrar_t cuda, void *opaque, uint8_t bitdepth)
{
  registrar_t next = (registrar_t)dlsym(RTLD_NEXT, sym);
  if (!next) { fprintf(stderr, "uvg266_b200: %s not found in the host library\n", sym); return 0; }
  int ok = next(opaque, bitdepth);
  if (ok && group_enabled(group)) ok = cuda(opaque, bitdepth);
  else if (ok && !strcmp(group, "picture") && group_enabled("hook")) ok = uvg_b200_install_rough_hook_only(bitdepth);
  return ok;
}

int uvg_strategy_register_picture(void *opaque, uint8_t bitdepth)
{ return forward("uvg_strategy_register_picture", "picture", uvg_strategy_register_picture_cuda, opaque, bitdepth); }
int uvg_strategy_register_dct(void *opaque, uint8_t bitdepth)
{ return forward("uvg_strategy_register_dct", "dct", uvg_strategy_register_dct_cuda, opaque, bitdepth); }
int uvg_strategy_register_quant(void *opaque, uint8_t bitdepth)
{ return forward("uvg_strategy_register_quant", "quant", uvg_strategy_register_quant_cuda, opaque, bitdepth); }
int uvg_strategy_register_ipol(void *opaque, uint8_t bitdepth)
{ return forward("uvg_strategy_register_ipol", "ipol", uvg_strategy_register_ipol_cuda, opaque, bitdepth); }
int uvg_strategy_register_intra(void *opaque, uint8_t bitdepth)
{ return forward("uvg_strategy_register_intra", "intra", uvg_strategy_register_intra_cuda, opaque, bitdepth); }
