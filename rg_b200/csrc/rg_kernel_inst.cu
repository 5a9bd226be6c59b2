// rg_kernel_inst.cu -- one instantiation of the fused kernel: compiled once per (RG_INST_J, RG_INST_BIG) by the Makefile.
// Exports the kernel's address and the upload of a plan into THIS translation unit's constant-memory plan table.
#include <cuda_runtime.h>
#ifndef RG_INST_J
#error "compile with -DRG_INST_J=<Taylor order> -DRG_INST_BIG=<0|1>"
#endif
#define RG_CAT2(a, j, b) a##j##_##b
#define RG_CAT(a, j, b) RG_CAT2(a, j, b)
#define RG_PLAN_TABLE RG_CAT(c_plan_, RG_INST_J, RG_INST_BIG)
#include "rg_kernel.cuh"

extern "C" const void* RG_CAT(rg_inst_kernel_, RG_INST_J, RG_INST_BIG)(void) {
    return (const void*)&rgdev::rg_kernel<RG_INST_J, (RG_INST_BIG != 0)>;
}
extern "C" int RG_CAT(rg_inst_upload_, RG_INST_J, RG_INST_BIG)(const rgdev::DevPlan* dp, int slot) {
    return (int)cudaMemcpyToSymbol(rgdev::c_plan, dp, sizeof(rgdev::DevPlan), sizeof(rgdev::DevPlan) * (size_t)slot);
}
