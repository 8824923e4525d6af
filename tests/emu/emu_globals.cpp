// Host-emulation globals (tests only; see world_b200/csrc/wb_platform.cuh).
#include "../../world_b200/csrc/wb_platform.cuh"
wb_dim3 threadIdx, blockIdx, blockDim, gridDim;
alignas(64) unsigned char wb_emu_smem[WB_EMU_SMEM_BYTES];
