// trie.cu -- builders: M (mptize), S (state root), U (resident trie).  PLACEHOLDER: not implemented yet.
#include "../../include/phant_gpu.h"
#include "ctx.cuh"
extern "C" int phant_gpu_mpt_root(phant_gpu_ctx*, const uint8_t*, const uint32_t*, const uint8_t*, const uint64_t*, uint64_t, uint8_t*) { return PHANT_GPU_E_INVALID; }
extern "C" int phant_gpu_state_root(phant_gpu_ctx*, const phant_gpu_accounts*, uint8_t*) { return PHANT_GPU_E_INVALID; }
extern "C" int phant_gpu_trie_open(phant_gpu_ctx*, const phant_gpu_trie_desc*, phant_gpu_trie**) { return PHANT_GPU_E_INVALID; }
extern "C" int phant_gpu_trie_root(phant_gpu_trie*, uint8_t*) { return PHANT_GPU_E_INVALID; }
extern "C" int phant_gpu_trie_update(phant_gpu_trie*, const uint8_t*, const uint8_t*, const uint32_t*, uint64_t, uint8_t*) { return PHANT_GPU_E_INVALID; }
extern "C" void phant_gpu_trie_close(phant_gpu_trie*) {}
