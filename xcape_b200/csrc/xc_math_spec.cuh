// xc_math_spec.cuh — the "SPEC" transcendentals (DESIGN.md §SPEC math), device side.
//
// binary32 in / binary32 out, binary64 inside.  Only IEEE-754 round-to-nearest +, -, *,
// fma and integer operations are used, so the same sequence of roundings can be (and is,
// independently, in oracle/xcape_oracle.cpp) reproduced on a CPU: results are bit-identical
// across the two by construction.  Against the correctly-rounded binary32 function they
// differ with probability ~2^-20 per call (internal relative error < 2^-44), i.e. they are
// "valid libm" replacements for gfortran's expf/logf/powf in the reference
// (CAPE_CODE_model_lev.f90:236,438-462,570-620).
//
// B200 note: FP64 runs at half the FP32 rate on sm_100a, which is what makes a
// double-precision core cheaper here than a float-float one.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace xc {

// Polynomial / reduction constants live in constant memory: DFMA takes a c[bank][offset]
// operand directly, whereas an immediate binary64 costs two UMOV issue slots per use
// (ncu r1a: 40 of the 195 instructions of the moist-iteration body were such UMOVs).
__constant__ double kExp[16] = {
    0x1.71547652b82fep+0,    // [0]  log2(e)
    0x1.62e42fee00000p-1,    // [1]  ln2 hi
    0x1.a39ef35793c76p-33,   // [2]  ln2 lo
    0x1.ae64567f544e4p-26,   // [3]  1/11!
    0x1.27e4fb7789f5cp-22,   // [4]  1/10!
    0x1.71de3a556c734p-19,   // [5]  1/9!
    0x1.a01a01a01a01ap-16,   // [6]  1/8!
    0x1.a01a01a01a01ap-13,   // [7]  1/7!
    0x1.6c16c16c16c17p-10,   // [8]  1/6!
    0x1.1111111111111p-7,    // [9]  1/5!
    0x1.5555555555555p-5,    // [10] 1/4!
    0x1.5555555555555p-3,    // [11] 1/3!
    0.5, 1.0, 6755399441055744.0, -6755399441055744.0};
__constant__ double kLog[8] = {
    0x1.2492492492492p-3,    // [0]  1/7
    -0x1.5555555555555p-3,   // [1] -1/6
    0x1.999999999999ap-3,    // [2]  1/5
    0x1.5555555555555p-2,    // [3]  1/3
    0x1.62e42fee00000p-1,    // [4] ln2 hi (32 significant bits)
    0x1.a39ef35793c76p-33,   // [5] ln2 lo
    0.0, 0.0};
// SPEC log table: z in [0.6875, 1.375) in 128 intervals (2^-8 wide below 1, 2^-7 above); .x = invc ~ 1/c
// (c = interval centre; c = 1 for the two intervals touching 1), .y = logc = log(1/invc) for the rounded invc.
// Generated with 80-digit arithmetic; identical to SP_LOG_T in oracle/xcape_oracle.cpp.
__device__ const double2 kLogT[128] = {
    {0x1.734f0c541fe8dp+0, -0x1.7cc7f7db46a0ep-2},
    {0x1.713786d9c7c09p+0, -0x1.76feecb947176p-2},
    {0x1.6f26016f26017p+0, -0x1.713e33a46a17cp-2},
    {0x1.6d1a62681c861p+0, -0x1.6b85b4cffa3fdp-2},
    {0x1.6b1490aa31a3dp+0, -0x1.65d558d4ce00bp-2},
    {0x1.691473a88d0c0p+0, -0x1.602d08af091ecp-2},
    {0x1.6719f3601671ap+0, -0x1.5a8cadbbedfa1p-2},
    {0x1.6524f853b4aa3p+0, -0x1.54f431b7be1a8p-2},
    {0x1.63356b88ac0dep+0, -0x1.4f637ebba9810p-2},
    {0x1.614b36831ae94p+0, -0x1.49da7f3bcc420p-2},
    {0x1.5f66434292dfcp+0, -0x1.44591e0539f49p-2},
    {0x1.5d867c3ece2a5p+0, -0x1.3edf463c1683ep-2},
    {0x1.5babcc647fa91p+0, -0x1.396ce359bbf53p-2},
    {0x1.59d61f123ccaap+0, -0x1.3401e12aecba0p-2},
    {0x1.5805601580560p+0, -0x1.2e9e2bce12286p-2},
    {0x1.56397ba7c52e2p+0, -0x1.2941afb186b7cp-2},
    {0x1.54725e6bb82fep+0, -0x1.23ec5991eba49p-2},
    {0x1.52aff56a8054bp+0, -0x1.1e9e1678899f5p-2},
    {0x1.50f22e111c4c5p+0, -0x1.1956d3b9bc2f9p-2},
    {0x1.4f38f62dd4c9bp+0, -0x1.14167ef367784p-2},
    {0x1.4d843bedc2c4cp+0, -0x1.0edd060b78082p-2},
    {0x1.4bd3edda68fe1p+0, -0x1.09aa572e6c6d4p-2},
    {0x1.4a27fad76014ap+0, -0x1.047e60cde83b7p-2},
    {0x1.4880522014880p+0, -0x1.feb2233ea07cbp-3},
    {0x1.46dce34596066p+0, -0x1.f474b134df228p-3},
    {0x1.453d9e2c776cap+0, -0x1.ea4449f04aaf5p-3},
    {0x1.43a2730abee4dp+0, -0x1.e020cc6235ab5p-3},
    {0x1.420b5265e5951p+0, -0x1.d60a17f903514p-3},
    {0x1.40782d10e6566p+0, -0x1.cc000c9db3c52p-3},
    {0x1.3ee8f42a5af07p+0, -0x1.c2028ab17f9b5p-3},
    {0x1.3d5d991aa75c6p+0, -0x1.b811730b823d4p-3},
    {0x1.3bd60d9232955p+0, -0x1.ae2ca6f672bd8p-3},
    {0x1.3a524387ac822p+0, -0x1.a454082e6ab03p-3},
    {0x1.38d22d366088ep+0, -0x1.9a8778debaa3ap-3},
    {0x1.3755bd1c945eep+0, -0x1.90c6db9fcbcdbp-3},
    {0x1.35dce5f9f2af8p+0, -0x1.871213750e994p-3},
    {0x1.34679ace01346p+0, -0x1.7d6903caf5acdp-3},
    {0x1.32f5ced6a1dfap+0, -0x1.73cb9074fd14dp-3},
    {0x1.3187758e9ebb6p+0, -0x1.6a399dabbd383p-3},
    {0x1.301c82ac40260p+0, -0x1.60b3100b09474p-3},
    {0x1.2eb4ea1fed14bp+0, -0x1.5737cc9018cddp-3},
    {0x1.2d50a012d50a0p+0, -0x1.4dc7b897bc1c7p-3},
    {0x1.2bef98e5a3711p+0, -0x1.4462b9dc9b3dcp-3},
    {0x1.2a91c92f3c105p+0, -0x1.3b08b6757f2a7p-3},
    {0x1.293725bb804a5p+0, -0x1.31b994d3a4f86p-3},
    {0x1.27dfa38a1ce4dp+0, -0x1.28753bc11aba2p-3},
    {0x1.268b37cd60127p+0, -0x1.1f3b925f25d44p-3},
    {0x1.2539d7e9177b2p+0, -0x1.160c8024b27b0p-3},
    {0x1.23eb79717605bp+0, -0x1.0ce7ecdccc28bp-3},
    {0x1.22a0122a0122ap+0, -0x1.03cdc0a51ec0dp-3},
    {0x1.21579804855e6p+0, -0x1.f57bc7d9005dbp-4},
    {0x1.2012012012012p+0, -0x1.e3707ee30487bp-4},
    {0x1.1ecf43c7fb84cp+0, -0x1.d179788219362p-4},
    {0x1.1d8f5672e4abdp+0, -0x1.bf968769fca18p-4},
    {0x1.1c522fc1ce059p+0, -0x1.adc77ee5aea8ep-4},
    {0x1.1b17c67f2bae3p+0, -0x1.9c0c32d4d254dp-4},
    {0x1.19e0119e0119ep+0, -0x1.8a6477a91dc29p-4},
    {0x1.18ab083902bdbp+0, -0x1.78d02263d82d7p-4},
    {0x1.1778a191bd684p+0, -0x1.674f089365a78p-4},
    {0x1.1648d50fc3201p+0, -0x1.55e10050e0382p-4},
    {0x1.151b9a3fdd5c9p+0, -0x1.4485e03dbdfb0p-4},
    {0x1.13f0e8d344724p+0, -0x1.333d7f8183f4ap-4},
    {0x1.12c8b89edc0acp+0, -0x1.2207b5c7854a1p-4},
    {0x1.11a3019a74826p+0, -0x1.10e45b3cae829p-4},
    {0x1.107fbbe011080p+0, -0x1.ffa6911ab9309p-5},
    {0x1.0f5edfab325a2p+0, -0x1.dda8adc67ee59p-5},
    {0x1.0e40655826011p+0, -0x1.bbcebfc68f424p-5},
    {0x1.0d24456359e3ap+0, -0x1.9a187b573de81p-5},
    {0x1.0c0a7868b4171p+0, -0x1.788595a3577c8p-5},
    {0x1.0af2f722eecb5p+0, -0x1.5715c4c03cee1p-5},
    {0x1.09ddba6af8360p+0, -0x1.35c8bfaa13069p-5},
    {0x1.08cabb37565e2p+0, -0x1.149e3e4005a8dp-5},
    {0x1.07b9f29b8eae2p+0, -0x1.e72bf2813ce6ap-6},
    {0x1.06ab59c7912fbp+0, -0x1.a55f548c5c427p-6},
    {0x1.059eea0727586p+0, -0x1.63d6178690bbep-6},
    {0x1.04949cc1664c5p+0, -0x1.228fb1fea2e0ap-6},
    {0x1.038c6b78247fcp+0, -0x1.c317384c75f0dp-7},
    {0x1.02864fc7729e9p+0, -0x1.41929f968330cp-7},
    {0x1.0182436517a37p+0, -0x1.8121214586b02p-8},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.fa11caa01fa12p-1, 0x1.7dc475f810a69p-7},
    {0x1.f6310aca0dbb5p-1, 0x1.3cea44346a584p-6},
    {0x1.f25f644230ab5p-1, 0x1.b9fc027af919ap-6},
    {0x1.ee9c7f8458e02p-1, 0x1.1b0d98923d97fp-5},
    {0x1.eae807aba01ebp-1, 0x1.58a5bafc8e4d3p-5},
    {0x1.e741aa59750e4p-1, 0x1.95c830ec8e3f2p-5},
    {0x1.e3a9179dc1a73p-1, 0x1.d276b8adb0b56p-5},
    {0x1.e01e01e01e01ep-1, 0x1.075983598e471p-4},
    {0x1.dca01dca01dcap-1, 0x1.253f62f0a1417p-4},
    {0x1.d92f2231e7f8ap-1, 0x1.42edcbea646eep-4},
    {0x1.d5cac807572b2p-1, 0x1.60658a93750c4p-4},
    {0x1.d272ca3fc5b1ap-1, 0x1.7da766d7b12d0p-4},
    {0x1.cf26e5c44bfc6p-1, 0x1.9ab42462033aep-4},
    {0x1.cbe6d9601cbe7p-1, 0x1.b78c82bb0eda0p-4},
    {0x1.c8b265afb8a42p-1, 0x1.d4313d66cb35dp-4},
    {0x1.c5894d10d4986p-1, 0x1.f0a30c01162a4p-4},
    {0x1.c26b5392ea01cp-1, 0x1.0671512ca596fp-3},
    {0x1.bf583ee868d8bp-1, 0x1.14785846742acp-3},
    {0x1.bc4fd65883e7bp-1, 0x1.2266f190a5acdp-3},
    {0x1.b951e2b18ff23p-1, 0x1.303d718e47fd5p-3},
    {0x1.b65e2e3beee05p-1, 0x1.3dfc2b0ecc62ap-3},
    {0x1.b37484ad806cep-1, 0x1.4ba36f39a55e5p-3},
    {0x1.b094b31d922a4p-1, 0x1.59338d9982085p-3},
    {0x1.adbe87f94905ep-1, 0x1.66acd4272ad51p-3},
    {0x1.aaf1d2f87ebfdp-1, 0x1.740f8f54037a3p-3},
    {0x1.a82e65130e159p-1, 0x1.815c0a14357e9p-3},
    {0x1.a574107688a4ap-1, 0x1.8e928de886d41p-3},
    {0x1.a2c2a87c51ca0p-1, 0x1.9bb362e7dfb85p-3},
    {0x1.a01a01a01a01ap-1, 0x1.a8becfc882f19p-3},
    {0x1.9d79f176b682dp-1, 0x1.b5b519e8fb5a6p-3},
    {0x1.9ae24ea5510dap-1, 0x1.c2968558c18c2p-3},
    {0x1.9852f0d8ec0ffp-1, 0x1.cf6354e09c5ddp-3},
    {0x1.95cbb0be377aep-1, 0x1.dc1bca0abec7bp-3},
    {0x1.934c67f9b2ce6p-1, 0x1.e8c0252aa5a60p-3},
    {0x1.90d4f120190d5p-1, 0x1.f550a564b7b37p-3},
    {0x1.8e6527af1373fp-1, 0x1.00e6c45ad501dp-2},
    {0x1.8bfce8062ff3ap-1, 0x1.071b85fcd590dp-2},
    {0x1.899c0f601899cp-1, 0x1.0d46b579ab74bp-2},
    {0x1.87427bcc092b9p-1, 0x1.136870293a8b0p-2},
    {0x1.84f00c2780614p-1, 0x1.1980d2dd4236fp-2},
    {0x1.82a4a0182a4a0p-1, 0x1.1f8ff9e48a2f3p-2},
    {0x1.8060180601806p-1, 0x1.2596010df763ap-2},
    {0x1.7e225515a4f1dp-1, 0x1.2b9303ab89d25p-2},
    {0x1.7beb3922e017cp-1, 0x1.31871c9544185p-2},
    {0x1.79baa6bb6398bp-1, 0x1.3772662bfd85cp-2},
    {0x1.77908119ac60dp-1, 0x1.3d54fa5c1f710p-2},
    {0x1.756cac201756dp-1, 0x1.432ef2a04e813p-2},
};

// 2^(j/128), j = 0..127, correctly rounded to binary64.  Read through the read-only data path
// (__ldg): 1 KB, L1-resident; the index differs per lane, which constant memory would serialise.
__device__ const double kExpT[128] = {
    0x1.0000000000000p+0, 0x1.0163da9fb3335p+0, 0x1.02c9a3e778061p+0, 0x1.04315e86e7f85p+0,
    0x1.059b0d3158574p+0, 0x1.0706b29ddf6dep+0, 0x1.0874518759bc8p+0, 0x1.09e3ecac6f383p+0,
    0x1.0b5586cf9890fp+0, 0x1.0cc922b7247f7p+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0fb66affed31bp+0,
    0x1.11301d0125b51p+0, 0x1.12abdc06c31ccp+0, 0x1.1429aaea92de0p+0, 0x1.15a98c8a58e51p+0,
    0x1.172b83c7d517bp+0, 0x1.18af9388c8deap+0, 0x1.1a35beb6fcb75p+0, 0x1.1bbe084045cd4p+0,
    0x1.1d4873168b9aap+0, 0x1.1ed5022fcd91dp+0, 0x1.2063b88628cd6p+0, 0x1.21f49917ddc96p+0,
    0x1.2387a6e756238p+0, 0x1.251ce4fb2a63fp+0, 0x1.26b4565e27cddp+0, 0x1.284dfe1f56381p+0,
    0x1.29e9df51fdee1p+0, 0x1.2b87fd0dad990p+0, 0x1.2d285a6e4030bp+0, 0x1.2ecafa93e2f56p+0,
    0x1.306fe0a31b715p+0, 0x1.32170fc4cd831p+0, 0x1.33c08b26416ffp+0, 0x1.356c55f929ff1p+0,
    0x1.371a7373aa9cbp+0, 0x1.38cae6d05d866p+0, 0x1.3a7db34e59ff7p+0, 0x1.3c32dc313a8e5p+0,
    0x1.3dea64c123422p+0, 0x1.3fa4504ac801cp+0, 0x1.4160a21f72e2ap+0, 0x1.431f5d950a897p+0,
    0x1.44e086061892dp+0, 0x1.46a41ed1d0057p+0, 0x1.486a2b5c13cd0p+0, 0x1.4a32af0d7d3dep+0,
    0x1.4bfdad5362a27p+0, 0x1.4dcb299fddd0dp+0, 0x1.4f9b2769d2ca7p+0, 0x1.516daa2cf6642p+0,
    0x1.5342b569d4f82p+0, 0x1.551a4ca5d920fp+0, 0x1.56f4736b527dap+0, 0x1.58d12d497c7fdp+0,
    0x1.5ab07dd485429p+0, 0x1.5c9268a5946b7p+0, 0x1.5e76f15ad2148p+0, 0x1.605e1b976dc09p+0,
    0x1.6247eb03a5585p+0, 0x1.6434634ccc320p+0, 0x1.6623882552225p+0, 0x1.68155d44ca973p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6c012750bdabfp+0, 0x1.6dfb23c651a2fp+0, 0x1.6ff7df9519484p+0,
    0x1.71f75e8ec5f74p+0, 0x1.73f9a48a58174p+0, 0x1.75feb564267c9p+0, 0x1.780694fde5d3fp+0,
    0x1.7a11473eb0187p+0, 0x1.7c1ed0130c132p+0, 0x1.7e2f336cf4e62p+0, 0x1.80427543e1a12p+0,
    0x1.82589994cce13p+0, 0x1.8471a4623c7adp+0, 0x1.868d99b4492edp+0, 0x1.88ac7d98a6699p+0,
    0x1.8ace5422aa0dbp+0, 0x1.8cf3216b5448cp+0, 0x1.8f1ae99157736p+0, 0x1.9145b0b91ffc6p+0,
    0x1.93737b0cdc5e5p+0, 0x1.95a44cbc8520fp+0, 0x1.97d829fde4e50p+0, 0x1.9a0f170ca07bap+0,
    0x1.9c49182a3f090p+0, 0x1.9e86319e32323p+0, 0x1.a0c667b5de565p+0, 0x1.a309bec4a2d33p+0,
    0x1.a5503b23e255dp+0, 0x1.a799e1330b358p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ac36bbfd3f37ap+0,
    0x1.ae89f995ad3adp+0, 0x1.b0e07298db666p+0, 0x1.b33a2b84f15fbp+0, 0x1.b59728de5593ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.ba5b030a1064ap+0, 0x1.bcc1e904bc1d2p+0, 0x1.bf2c25bd71e09p+0,
    0x1.c199bdd85529cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c67f12e57d14bp+0, 0x1.c8f6d9406e7b5p+0,
    0x1.cb720dcef9069p+0, 0x1.cdf0b555dc3fap+0, 0x1.d072d4a07897cp+0, 0x1.d2f87080d89f2p+0,
    0x1.d5818dcfba487p+0, 0x1.d80e316c98398p+0, 0x1.da9e603db3285p+0, 0x1.dd321f301b460p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e264614f5a129p+0, 0x1.e502ee78b3ff6p+0, 0x1.e7a51fbc74c83p+0,
    0x1.ea4afa2a490dap+0, 0x1.ecf482d8e67f1p+0, 0x1.efa1bee615a27p+0, 0x1.f252b376bba97p+0,
    0x1.f50765b6e4540p+0, 0x1.f7bfdad9cbe14p+0, 0x1.fa7c1819e90d8p+0, 0x1.fd3c22b8f71f1p+0};
__constant__ double kExp2[8] = {
    0x1.71547652b82fep+7,     // [0] 128 log2(e)
    0x1.62e42fee00000p-8,     // [1] ln2/128 hi (32 significant bits)
    0x1.a39ef35793c76p-40,    // [2] ln2/128 lo
    0x1.5555555555555p-5,     // [3] 1/24
    0x1.5555555555555p-3,     // [4] 1/6
    0.5, 6755399441055744.0, -6755399441055744.0};

// exp of a double already known to lie in [-700, 700]; returns double.
//   n = rint(128 x log2 e), r = x - n ln2/128 (|r| <= 0.0028), j = n mod 128, e = floor(n/128)
//   exp(x) = 2^e * T[j] * (1 + r(1 + r(1/2 + r(1/6 + r/24))))          truncation < 2^-49
// 10 FP64 issue slots against 16 for the table-free degree-11 form it replaces (ncu r1c: FP64 and
// XU instructions, at 2+ dispatch cycles each, cap the faithful kernel's IPC).
__device__ __forceinline__ double spec_exp_core(double x) {
  const double tm = __dadd_rn(__dmul_rn(x, kExp2[0]), kExp2[6]);
  const double nd = __dsub_rn(tm, kExp2[6]);   // same value as tm + (-magic): one constant fewer to load
  const int n = __double2loint(tm);          // low word of t+1.5*2^52 is rint(t) in two's complement
  double r = __fma_rn(-nd, kExp2[1], x);
  r = __fma_rn(-nd, kExp2[2], r);
  double q = kExp2[3];
  q = __fma_rn(q, r, kExp2[4]);
  q = __fma_rn(q, r, 0.5);                      // literal: fits the instruction's 32-bit immediate
  q = __fma_rn(q, r, 1.0);
  q = __fma_rn(q, r, 1.0);
  const double s = __dmul_rn(__ldg(&kExpT[n & 127]), q);
  return __hiloint2double(__double2hiint(s) + ((n >> 7) << 20), __double2loint(s));
}

// exp of a double with |x| <= 2^-3: no range reduction (n == 0).  |x| <= 2^-6 (the usual case for
// theta2 = theta1*exp(.), f90:460-462): Taylor degree 5 (truncation < 2^-45); else degree 8.
__device__ __forceinline__ double spec_exp_small_core(double x, bool tiny) {
  double p;
  if (tiny) {
    p = kExp[9];                             // 1/5!
  } else {
    p = kExp[6];                             // 1/8!
    p = __fma_rn(p, x, kExp[7]);
    p = __fma_rn(p, x, kExp[8]);
    p = __fma_rn(p, x, kExp[9]);
  }
  p = __fma_rn(p, x, kExp[10]);
  p = __fma_rn(p, x, kExp[11]);
  p = __fma_rn(p, x, 0.5);
  p = __fma_rn(p, x, 1.0);
  p = __fma_rn(p, x, 1.0);
  return p;
}

__device__ __forceinline__ double spec_exp_d(double x) {
  if (!(x >= -700.0)) return (x != x) ? x : 0.0;
  if (x > 700.0) return CUDART_INF;
  return spec_exp_core(x);
}

// log of a positive, finite, normal double: x = 2^k z, r = fma(z, invc, -1) (|r| <= 2^-7),
//   log x = (k ln2_hi + logc) + (k ln2_lo + log1p(r)),  log1p by its degree-7 Taylor polynomial.
// 11 FP64 instructions against 24 for the table-free atanh form it replaces.
__device__ __forceinline__ double spec_log_core(double x) {
  const int hi32 = __double2hiint(x);
  const int tmp = hi32 - 0x3fe60000;               // the low word of the offset is zero: no borrow
  const int k = tmp >> 20;                          // arithmetic shift
  const double z = __hiloint2double(hi32 - (tmp & 0xfff00000), __double2loint(x));
  const double2 t = __ldg(&kLogT[(tmp >> 13) & 127]);
  const double r = __fma_rn(z, t.x, -1.0);
  const double kd = (double)k;
  double q = kLog[0];
  q = __fma_rn(q, r, kLog[1]);
  q = __fma_rn(q, r, kLog[2]);
  q = __fma_rn(q, r, -0.25);
  q = __fma_rn(q, r, kLog[3]);
  q = __fma_rn(q, r, -0.5);
  const double r2 = __dmul_rn(r, r);
  const double lp = __fma_rn(q, r2, r);
  const double hi = __fma_rn(kd, kLog[4], t.y);
  const double lo = __fma_rn(kd, kLog[5], lp);
  return __dadd_rn(hi, lo);
}

__device__ __forceinline__ double spec_log_d(double x) {
  if (!(x > 0.0)) return (x == 0.0) ? -CUDART_INF : CUDART_NAN;
  if (x == CUDART_INF) return x;
  return spec_log_core(x);
}

// ---- binary32 front ends -------------------------------------------------------------
// The range guards are evaluated on the binary32 argument (FP32 pipe) — equivalent to the
// binary64 guards of the spec because +-700 are exact in binary32.
__device__ __forceinline__ float spec_expf(float x) {
  if (fabsf(x) <= 700.0f) return __double2float_rn(spec_exp_core((double)x));   // one compare on the hot path
  return (x != x) ? x : (x < 0.0f ? 0.0f : CUDART_INF_F);
}
// exp for arguments that are almost always tiny (|x| <= 2^-3 takes the reduction-free path)
__device__ __forceinline__ float spec_expf_small(float x) {
  const float ax = fabsf(x);
  // three separate returns: written as one call with a `tiny` flag, nvcc if-converts the tiers and the
  // usual (tiny) case pays for the three extra DFMAs of the degree-8 tier plus two selects
  if (ax <= 0.015625f) return __double2float_rn(spec_exp_small_core((double)x, true));
  if (ax <= 0.125f) return __double2float_rn(spec_exp_small_core((double)x, false));
  return spec_expf(x);
}
// ---- binary32 exp ("SPEC exp32"): FP32 pipe only, no conversions, no FP64 ----------------------------
// The moist iteration evaluates exp twice per pass (Bolton's es(T), f90:570-581, and the theta update,
// f90:460-462).  With the binary64 core each call cost two F2F conversions (XU pipe, 8 cycles per warp
// instruction) and 10 resp. 5 FP64 instructions (2 cycles each); ncu r1u showed XU as the busiest pipe.
// This version stays in binary32 with float-float tails:
//   t = fma(x, 1024 log2e, 1.5*2^23); n = rint(..) sits in t's low mantissa bits; nf = t - 1.5*2^23
//   r = fma(nf, -L2, fma(nf, -L1, x))        L1 = RN32(ln2/1024) (the inner fma is exact), L2 = RN32(ln2/1024 - L1)
//   p = r + r^2 (1/2 + r/6)                   |r| <= 2^-11.5: the next term is < 2^-50
//   y = Th[j] + fma(Th[j], p, Tl[j])          j = n mod 1024, {Th, Tl} = 2^(j/1024) as a binary32 pair
//   exp(x) = y * 2^(n div 1024)               added to the exponent field
// Pre-rounding error ~2^-11 ulp: on 5e7 arguments in [-53.6, 7.1] it differs from the correctly rounded
// binary32 exp in 1.0e-4 of the calls, never by more than one ulp (glibc's expf: 6.2e-4) — a valid libm.
// Only IEEE binary32 +, -, *, fma and integer operations: oracle/xcape_oracle.cpp reproduces it bit for bit.
__device__ const float2 kExp32T[1024] = {
#include "xc_exp32_table.inc"
};
namespace e32 {
constexpr float kKL = 0x1.715476p+10f;     // RN32(1024 log2 e)
constexpr float kL1 = 0x1.62e43p-11f;      // RN32(ln2 / 1024)
constexpr float kL2 = -0x1.05c61p-39f;     // RN32(ln2 / 1024 - L1)
constexpr float kMagic = 12582912.0f;      // 1.5 * 2^23
constexpr float kC3 = 0x1.555556p-3f;      // RN32(1/6)
}  // namespace e32

// The CAPE kernels keep a copy of the table in shared memory (8 KB per CTA): the index differs per lane, and
// LDS with a scaled index costs two instructions where the global-memory path costs five (64-bit address
// arithmetic) plus an L1 round trip.  A kernel that uses the exp32 functions calls exp32_smem_fill() first.
#ifndef XC_EXP32_GLOBAL_TABLE
// one array, hi parts in [0, 1024) and lo parts in [1024, 2048): one address computation serves both loads (the second
// is the first plus an immediate offset), and the two-column kernel loads the four values of a pair straight into
// the halves of two packed registers
__shared__ float sExp32[2048];
__device__ __forceinline__ void exp32_smem_fill() {
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
    const float2 T = kExp32T[k];
    sExp32[k] = T.x; sExp32[1024 + k] = T.y;
  }
  __syncthreads();
}
__device__ __forceinline__ float exp32_hi(int j) { return sExp32[j]; }
__device__ __forceinline__ float exp32_lo(int j) { return sExp32[1024 + j]; }
#else
__device__ __forceinline__ void exp32_smem_fill() {}
__device__ __forceinline__ float exp32_hi(int j) { return __ldg(&kExp32T[j].x); }
__device__ __forceinline__ float exp32_lo(int j) { return __ldg(&kExp32T[j].y); }
#endif
__device__ __forceinline__ float2 exp32_entry(int j) { return make_float2(exp32_hi(j), exp32_lo(j)); }

// y in [1, 2) and the integer n (n div 1024 = binary exponent) for |x| <= 104
__device__ __forceinline__ float spec32_exp_mant(float x, int& n) {
  const float t = __fmaf_rn(x, e32::kKL, e32::kMagic);
  const float nf = __fsub_rn(t, e32::kMagic);
  float r = __fmaf_rn(nf, -e32::kL1, x);
  r = __fmaf_rn(nf, -e32::kL2, r);
  const int bt = __float_as_int(t);
  const float2 T = exp32_entry(bt & 1023);
  const float q = __fmaf_rn(r, e32::kC3, 0.5f);
  const float v = __fmul_rn(r, r);
  const float p = __fmaf_rn(q, v, r);
  const float s = __fmaf_rn(T.x, p, T.y);
  n = bt - 0x4B400000;
  return __fadd_rn(T.x, s);
}
// |x| <= 87 (or 0 < x <= 88.7228): the result is a normal binary32
__device__ __forceinline__ float spec32_exp_core(float x) {
  int n;
  const float y = spec32_exp_mant(x, n);
  return __int_as_float((int)(((unsigned)n >> 10) * 0x00800000u + (unsigned)__float_as_int(y)));   // SHF + LEA
}
__device__ __forceinline__ float spec32_expf(float x) {
  if (fabsf(x) <= 87.0f) return spec32_exp_core(x);
  if (x != x) return x;
  if (x > 88.72283172607421875f) return CUDART_INF_F;       // largest binary32 whose exp is finite
  if (x > 0.0f) return spec32_exp_core(x);                   // 2^127 y still fits
  if (x < -104.0f) return 0.0f;                              // exp(x) < 2^-150
  int n;                                                      // subnormal results: one rounding, in the last product
  const float y = spec32_exp_mant(x, n);
  const float big = __int_as_float(__float_as_int(y) + (((n >> 10) + 64) << 23));
  return __fmul_rn(big, 0x1p-64f);
}
// exp(x) for |x| <= 2^-6 (the theta update's argument): 1 + x is split exactly into h + e, the Taylor
// tail x^2 (1/2 + x/6 + x^2/24 + x^3/120) joins e, one final rounding.  != correctly rounded in 1.2e-5 of calls.
__device__ __forceinline__ float spec32_exp_tiny(float x) {
  const float h = __fadd_rn(1.0f, x);
  const float e = __fsub_rn(x, __fsub_rn(h, 1.0f));
  const float v = __fmul_rn(x, x);
  float u = __fmaf_rn(x, 0x1.111112p-7f, 0x1.555556p-5f);   // 1/120, 1/24
  u = __fmaf_rn(u, x, e32::kC3);
  u = __fmaf_rn(u, x, 0.5f);
  return __fadd_rn(h, __fmaf_rn(v, u, e));
}
__device__ __forceinline__ float spec32_expf_small(float x) {
  if (fabsf(x) <= 0.015625f) return spec32_exp_tiny(x);
  return spec32_expf(x);
}

__device__ __forceinline__ float spec_logf(float x) {
  if (!(x > 0.0f)) return (x == 0.0f) ? -CUDART_INF_F : CUDART_NAN_F;
  if (x == CUDART_INF_F) return x;
  return __double2float_rn(spec_log_core((double)x));
}
__device__ __forceinline__ float spec_powf(float x, float y) {
  return __double2float_rn(spec_exp_d(__dmul_rn((double)y, spec_log_d((double)x))));
}

}  // namespace xc
