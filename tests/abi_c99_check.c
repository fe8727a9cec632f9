/* abi_c99_check.c -- the header must be plain C (phant binds it with @cImport; no C++ in the signatures) and the
 * library must link from C.  Built and run by tests/test_abi.py::test_header_is_plain_c_and_links_from_c on the CPU box:
 * without a CUDA device create() has to fail with PHANT_GPU_E_NO_DEVICE, never fall back. */
#include "../include/phant_gpu.h"
#include <stdio.h>
#include <string.h>

int main(void)
{
    phant_gpu_ctx* ctx = NULL;
    phant_gpu_config cfg;
    phant_gpu_proof_batch batch;
    phant_gpu_witness witness;
    phant_gpu_accounts accounts;
    phant_gpu_trie_desc desc;
    phant_gpu_stats stats;
    int rc;
    memset(&cfg, 0, sizeof cfg);
    memset(&batch, 0, sizeof batch);
    memset(&witness, 0, sizeof witness);
    memset(&accounts, 0, sizeof accounts);
    memset(&desc, 0, sizeof desc);
    memset(&stats, 0, sizeof stats);
    if (phant_gpu_abi_version() != PHANT_GPU_ABI_VERSION) return 2;
    rc = phant_gpu_create(&ctx, &cfg);
    printf("create rc=%d (%s)\n", rc, phant_gpu_strerror(rc));
    if (rc == PHANT_GPU_OK) { /* a GPU is present: exercise one call and leave */
        unsigned char out[32];
        unsigned long long off[2] = {0, 0};
        rc = phant_gpu_keccak256_batch(ctx, (const uint8_t*)"", (const uint64_t*)off, 1, out);
        printf("keccak('') rc=%d first byte %02x\n", rc, out[0]);
        phant_gpu_destroy(ctx);
        return rc == 0 && out[0] == 0xc5 ? 0 : 3;
    }
    /* every entry point must reject a null context instead of crashing */
    if (phant_gpu_keccak256_batch(NULL, NULL, NULL, 1, NULL) != PHANT_GPU_E_INVALID) return 4;
    if (phant_gpu_verify_proofs(NULL, &batch, NULL, NULL, NULL, NULL) != PHANT_GPU_E_INVALID) return 5;
    if (phant_gpu_verify_witness(NULL, &witness, NULL, NULL, NULL, NULL) != PHANT_GPU_E_INVALID) return 6;
    if (phant_gpu_state_root(NULL, &accounts, NULL) != PHANT_GPU_E_INVALID) return 7;
    if (phant_gpu_mpt_root(NULL, NULL, NULL, NULL, NULL, 0, NULL) != PHANT_GPU_E_INVALID) return 8;
    if (phant_gpu_get_stats(NULL, &stats) != PHANT_GPU_E_INVALID) return 9;
    phant_gpu_destroy(NULL);
    return rc == PHANT_GPU_E_NO_DEVICE ? 0 : 10;
}
