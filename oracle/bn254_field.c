/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Instantiates the 4x64 Montgomery field template for BN254 Fr and Fq.
 * Follows halo2curves 0.1.0 src/bn256/fr.rs / fq.rs (pin Cargo.lock:1911-1913):
 * MODULUS, INV, R, R2, ROOT_OF_UNITY (=7^((r-1)/2^28)), ZETA (=7^(2(r-1)/3)), S=28.
 * The constants below are Montgomery-form limbs computed from those definitions and
 * cross-checked in tests against release-v0.13.1/chunk.protocol (domain.gen == ROOT^8,
 * n_inv == 2^-25) and release-v0.13.1/evm_verifier.yul:17-18 (moduli).
 */
#include "bn254_oracle.h"

const fr_t fr_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
const fr_t fr_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
const fr_t fr_ROOT_OF_UNITY = {{0x9632c7c5b639feb8ULL, 0x985ce3400d0ff299ULL, 0xb2dd880001b0ecd8ULL, 0x1d69070d6d98ce29ULL}};
const fr_t fr_ZETA = {{0x0363f29955fcd653ULL, 0x73e7950b5fc1e200ULL, 0xc5fce83e576d9d24ULL, 0x059c805da1c3a4d4ULL}};
const fr_t fr_GENERATOR = {{0x3057819e4fffffdbULL, 0x307f6d866832bb01ULL, 0x5c65ec9f484e3a89ULL, 0x0180a96573d3d9f8ULL}};

const fq_t fq_ONE = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}};
const fq_t fq_R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};

#define FP_NAME fr
#define FP_MOD0 0x43e1f593f0000001ULL
#define FP_MOD1 0x2833e84879b97091ULL
#define FP_MOD2 0xb85045b68181585dULL
#define FP_MOD3 0x30644e72e131a029ULL
#define FP_INV 0xc2e1f593efffffffULL
#include "fp_template.h"

#define FP_NAME fq
#define FP_MOD0 0x3c208c16d87cfd47ULL
#define FP_MOD1 0x97816a916871ca8dULL
#define FP_MOD2 0xb85045b68181585dULL
#define FP_MOD3 0x30644e72e131a029ULL
#define FP_INV 0x87d20782e4866389ULL
#include "fp_template.h"
