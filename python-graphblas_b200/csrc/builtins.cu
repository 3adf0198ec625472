// builtins.cu -- builtin types, operators, monoids, semirings, descriptors exported as data symbols
#include <map>
#include <string>

#include "grb_internal.h"

static GrB_Type_opaque type_objs[TC_COUNT] = {
    {TC_BOOL, 1, "GrB_BOOL"},     {TC_INT8, 1, "GrB_INT8"},     {TC_INT16, 2, "GrB_INT16"}, {TC_INT32, 4, "GrB_INT32"},
    {TC_INT64, 8, "GrB_INT64"},   {TC_UINT8, 1, "GrB_UINT8"},   {TC_UINT16, 2, "GrB_UINT16"},
    {TC_UINT32, 4, "GrB_UINT32"}, {TC_UINT64, 8, "GrB_UINT64"}, {TC_FP32, 4, "GrB_FP32"},   {TC_FP64, 8, "GrB_FP64"},
};
extern "C" {
GrB_Type GrB_BOOL = &type_objs[TC_BOOL], GrB_INT8 = &type_objs[TC_INT8], GrB_INT16 = &type_objs[TC_INT16],
         GrB_INT32 = &type_objs[TC_INT32], GrB_INT64 = &type_objs[TC_INT64], GrB_UINT8 = &type_objs[TC_UINT8],
         GrB_UINT16 = &type_objs[TC_UINT16], GrB_UINT32 = &type_objs[TC_UINT32], GrB_UINT64 = &type_objs[TC_UINT64],
         GrB_FP32 = &type_objs[TC_FP32], GrB_FP64 = &type_objs[TC_FP64];
}

const GrB_Type_opaque *type_of_code(int code) { return &type_objs[code]; }

#include "builtins_gen.inc"

static std::map<std::string, void *> *g_by_name = nullptr;

static void build_table() {
    if (g_by_name) return;
    g_by_name = new std::map<std::string, void *>();
    for (int i = 0; i < TC_COUNT; i++) (*g_by_name)[type_objs[i].name] = &type_objs[i];
    for (size_t i = 0; i < sizeof(g_symbol_table) / sizeof(g_symbol_table[0]); i++)
        (*g_by_name)[g_symbol_table[i].name] = g_symbol_table[i].handle;
}

void builtins_init() { build_table(); }

extern "C" void *GrB_cuda_lookup(const char *name) {
    build_table();
    if (!name) return nullptr;
    auto it = g_by_name->find(name);
    return it == g_by_name->end() ? nullptr : it->second;
}

extern "C" size_t GrB_cuda_symbol_names(char *buf, size_t buflen) {
    build_table();
    size_t need = 0;
    for (auto &kv : *g_by_name) need += kv.first.size() + 1;
    if (buf && buflen >= need) {
        char *p = buf;
        for (auto &kv : *g_by_name) {
            memcpy(p, kv.first.c_str(), kv.first.size() + 1);
            p += kv.first.size() + 1;
        }
    }
    return need;
}
