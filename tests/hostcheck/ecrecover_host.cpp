// Test-only: phant_b200/csrc/secp256k1.cuh (the per-thread device function of the ecrecover kernel) compiled as HOST code
// so that its arithmetic can be checked against the oracle and OpenSSL on a machine without a GPU.  Built as a shared
// object by tests/test_ecrecover_header_host.py; nothing in the product links or loads this.
#include <stdint.h>
#define __device__
#define __forceinline__ inline
#define __noinline__
#include "../../phant_b200/csrc/secp256k1.cuh"

extern "C" int host_ecrecover(const uint8_t* hash32, const uint8_t* sig65, uint8_t* pub65)
{
    pub65[0] = 0x04;
    return phant::secp::ecrecover(hash32, sig65, pub65 + 1) ? 1 : 0;
}
extern "C" void host_fp_mul(const uint8_t* a32, const uint8_t* b32, uint8_t* out32)
{
    using namespace phant::secp;
    to_be(out32, fp_mul(from_be(a32), from_be(b32)));
}
extern "C" void host_sc_mul(const uint8_t* a32, const uint8_t* b32, uint8_t* out32)
{
    using namespace phant::secp;
    to_be(out32, sc_mul(from_be(a32), from_be(b32)));
}
extern "C" void host_fp_addsub(const uint8_t* a32, const uint8_t* b32, uint8_t* sum32, uint8_t* diff32)
{
    using namespace phant::secp;
    to_be(sum32, fp_add(from_be(a32), from_be(b32)));
    to_be(diff32, fp_sub(from_be(a32), from_be(b32)));
}
