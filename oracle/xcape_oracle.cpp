// xcape_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch C++ restatement of the algorithms of the reference's eight
// f2py Fortran modules (xgcm/xcape v0.1.4, /root/reference/src/xcape/fortran).
// It exists so that the CUDA product path (xcape_b200/csrc) can be checked
// against the reference's arithmetic.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the
// product path never does.
//
// PARITY STATUS: pinned.  tests/test_oracle_golden.py checks this file against
// every golden vector the reference's own tests hold for the path
// (test/fixtures.py dataset_soundings + dataset_ERA5pressurelevel, committed
// as tests/golden/*.npz by tests/make_golden.py).  The reference itself cannot
// be compiled in this image (no Fortran compiler, no numpy.distutils/meson).
//
// Arithmetic contract (SURVEY.md §7 "Hard parts", Appendix A.8):
//   * CAPE chain: IEEE binary32 add/mul/div, no FMA contraction
//     (compile with -ffp-contract=off); ML accumulators binary64.
//   * transcendentals selectable per call (`tmode`):
//       0 = LIBM : glibc expf/logf/powf  (what gfortran's runtime links)
//       1 = CR   : (float)exp((double)x) etc. — correctly-rounded binary32
//       2 = SPEC : the deterministic double-precision algorithms of DESIGN.md
//                  §"SPEC math" (explicit fma; only IEEE +,-,*,fma and integer
//                  ops), which the CUDA kernel implements independently —
//                  oracle(SPEC) and the GPU agree bit for bit by construction.
//   * stdheight / SREH: binary64 with the reference's single-precision
//     literals; Bunkers: binary32.
//
// Each function cites the reference file:line it follows.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

#if defined(__GNUC__)
#define XC_FMA_TARGET __attribute__((target("fma")))
#else
#define XC_FMA_TARGET
#endif

namespace {

// ---------------------------------------------------------------------------
// SPEC math (DESIGN.md §SPEC math).  binary64 internals, binary32 in/out.
// ---------------------------------------------------------------------------
inline double bits2d(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; }
inline uint64_t d2bits(double d) { uint64_t b; std::memcpy(&b, &d, 8); return b; }

constexpr double SP_L2E    = 0x1.71547652b82fep+0;   // log2(e)
constexpr double SP_LN2_HI = 0x1.62e42fee00000p-1;   // fdlibm split of ln 2
constexpr double SP_LN2_LO = 0x1.a39ef35793c76p-33;
constexpr double SP_MAGIC  = 6755399441055744.0;     // 1.5 * 2^52

// 2^(j/128), j = 0..127, correctly rounded to binary64 (DESIGN.md "SPEC math")
static const double SP_EXP_T[128] = {
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
constexpr double SP_L2E_128    = 0x1.71547652b82fep+7;   // 128 log2(e)
constexpr double SP_LN2_128_HI = 0x1.62e42fee00000p-8;   // ln2/128, 32 significant bits
constexpr double SP_LN2_128_LO = 0x1.a39ef35793c76p-40;

XC_FMA_TARGET inline double spec_exp_d(double x) {
  if (!(x >= -700.0)) return (x != x) ? x : 0.0;
  if (x > 700.0) return INFINITY;
  double tm = x * SP_L2E_128 + SP_MAGIC;
  double nd = tm - SP_MAGIC;                            // rint(128 x log2 e) (RN mode, no re-association)
  double r = __builtin_fma(-nd, SP_LN2_128_HI, x);
  r = __builtin_fma(-nd, SP_LN2_128_LO, r);
  double q = 0x1.5555555555555p-5;                      // 1/24
  q = __builtin_fma(q, r, 0x1.5555555555555p-3);        // 1/6
  q = __builtin_fma(q, r, 0.5);
  q = __builtin_fma(q, r, 1.0);
  q = __builtin_fma(q, r, 1.0);
  int64_t n = (int64_t)nd;
  double s = SP_EXP_T[n & 127] * q;
  return bits2d(d2bits(s) + ((uint64_t)(n >> 7) << 52));
}

// natural log of a positive, finite, NORMAL binary64 (every positive finite binary32 converts to one).
// Table-driven: x = 2^k z with z in [0.6875, 1.375) cut into 128 intervals (2^-8 wide below 1, 2^-7 above);
// per interval invc ~ 1/c (c = centre; c = 1 for the two intervals that touch 1, so that nothing cancels near
// log(x) = 0) and logc = log(1/invc) for that ROUNDED invc, both correctly rounded from 80-digit arithmetic;
//   r = fma(z, invc, -1)  (|r| <= 2^-7),  log1p(r) by its degree-7 Taylor polynomial (truncation < 2^-52 |r|),
//   log x = (k ln2_hi + logc) + (k ln2_lo + log1p(r)).
struct SpLogTab { double invc, logc; };
static const SpLogTab SP_LOG_T[128] = {
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

XC_FMA_TARGET inline double spec_log_d(double x) {
  if (!(x > 0.0)) return (x == 0.0) ? -INFINITY : NAN;
  if (x == INFINITY) return x;
  const uint64_t ix = d2bits(x);
  const uint64_t tmp = ix - 0x3fe6000000000000ULL;
  const int i = (int)((tmp >> 45) & 127);
  const int64_t k = (int64_t)tmp >> 52;                                     // arithmetic shift
  const double z = bits2d(ix - (tmp & 0xfff0000000000000ULL));
  const double r = __builtin_fma(z, SP_LOG_T[i].invc, -1.0);
  const double kd = (double)k;
  double q = 0x1.2492492492492p-3;                  //  1/7
  q = __builtin_fma(q, r, -0x1.5555555555555p-3);   // -1/6
  q = __builtin_fma(q, r, 0x1.999999999999ap-3);    //  1/5
  q = __builtin_fma(q, r, -0.25);
  q = __builtin_fma(q, r, 0x1.5555555555555p-2);    //  1/3
  q = __builtin_fma(q, r, -0.5);
  const double r2 = r * r;
  const double lp = __builtin_fma(q, r2, r);        // log1p(r)
  const double hi = __builtin_fma(kd, SP_LN2_HI, SP_LOG_T[i].logc);
  const double lo = __builtin_fma(kd, SP_LN2_LO, lp);
  return hi + lo;
}

// ---------------------------------------------------------------------------
// SPEC exp32: the binary32 exp of SPEC mode (DESIGN.md §SPEC math).  Float-float tails, only IEEE
// binary32 +, -, *, fma and integer operations; written here independently of the CUDA version
// (xcape_b200/csrc/xc_math_spec.cuh), with which it agrees bit for bit by construction.
//   t = fma(x, 1024 log2e, 1.5*2^23)  ->  n = rint(1024 x log2e) in t's low mantissa bits
//   r = x - n ln2/1024 (two fmas, the first exact);  p = r + r^2 (1/2 + r/6)
//   exp(x) = 2^(n div 1024) * (Th + fma(Th, p, Tl)),  {Th, Tl} = 2^((n mod 1024)/1024) as a binary32 pair
// ---------------------------------------------------------------------------
struct SpF2 { float hi, lo; };
static const SpF2 SP_EXP32_T[1024] = {
#include "xc_exp32_table.inc"
};
inline int32_t f2bits(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline float bits2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

XC_FMA_TARGET inline float spec32_exp_mant(float x, int32_t& n) {
  const float t = __builtin_fmaf(x, 0x1.715476p+10f, 12582912.0f);
  const float nf = t - 12582912.0f;
  float r = __builtin_fmaf(nf, -0x1.62e43p-11f, x);
  r = __builtin_fmaf(nf, 0x1.05c61p-39f, r);
  const int32_t bt = f2bits(t);
  const SpF2 T = SP_EXP32_T[bt & 1023];
  const float q = __builtin_fmaf(r, 0x1.555556p-3f, 0.5f);
  const float v = r * r;
  const float p = __builtin_fmaf(q, v, r);
  const float s = __builtin_fmaf(T.hi, p, T.lo);
  n = bt - 0x4B400000;
  return T.hi + s;
}
XC_FMA_TARGET inline float spec32_expf(float x) {
  if (x != x) return x;
  if (x > 88.72283172607421875f) return INFINITY;             // largest binary32 whose exp is finite
  if (x < -104.0f) return 0.0f;
  int32_t n;
  const float y = spec32_exp_mant(x, n);
  const int32_t e = n >> 10;                                   // floor division
  if (x >= -87.0f) return bits2f(f2bits(y) + e * (1 << 23));   // normal result (e >= -126)
  const float big = bits2f(f2bits(y) + (e + 64) * (1 << 23));  // subnormal range: one rounding, in the last product
  return big * 0x1p-64f;
}
// |x| <= 2^-6: 1 + x = h + e exactly, Taylor tail joins e, one final rounding
XC_FMA_TARGET inline float spec32_exp_tiny(float x) {
  const float h = 1.0f + x;
  const float e = x - (h - 1.0f);
  const float v = x * x;
  float u = __builtin_fmaf(x, 0x1.111112p-7f, 0x1.555556p-5f);
  u = __builtin_fmaf(u, x, 0x1.555556p-3f);
  u = __builtin_fmaf(u, x, 0.5f);
  return h + __builtin_fmaf(v, u, e);
}

enum { T_LIBM = 0, T_CR = 1, T_SPEC = 2 };

template <int TM> inline float t_exp(float x) {
  if (TM == T_LIBM) return expf(x);
  if (TM == T_CR) return (float)exp((double)x);
  return spec32_expf(x);
}
// exp for the theta2 update (f90:460-462), whose argument is almost always tiny
template <int TM> inline float t_exp_small(float x) {
  if (TM == T_SPEC && std::fabs(x) <= 0.015625f) return spec32_exp_tiny(x);
  return t_exp<TM>(x);
}
template <int TM> inline float t_log(float x) {
  if (TM == T_LIBM) return logf(x);
  if (TM == T_CR) return (float)log((double)x);
  return (float)spec_log_d((double)x);
}
template <int TM> inline float t_pow(float x, float y) {
  if (TM == T_LIBM) return powf(x, y);
  if (TM == T_CR) return (float)pow((double)x, (double)y);
  return (float)spec_exp_d((double)y * spec_log_d((double)x));
}

inline float fmin_(float a, float b) { return (a < b) ? a : b; }   // Fortran MIN/MAX on NaN-free data
inline float fmax_(float a, float b) { return (a > b) ? a : b; }

// ---------------------------------------------------------------------------
// CAPE — getcape_ml == getcape_pl  (CAPE_CODE_model_lev.f90:97-564,
// CAPE_CODE_pressure_lev.f90:174-642; helper functions :570-620)
// ---------------------------------------------------------------------------
// constants: CAPE_CODE_model_lev.f90:188-211, every derived one folded in binary32
constexpr float c_g = 9.81f, c_p00 = 100000.0f, c_cp = 1005.7f, c_rd = 287.04f, c_rv = 461.5f;
constexpr float c_xlv = 2501000.0f, c_xls = 2836017.0f, c_t0 = 273.15f;
constexpr float c_cpv = 1875.0f, c_cpl = 4190.0f, c_cpi = 2118.636f;
constexpr float c_lv1 = c_xlv + (c_cpl - c_cpv) * c_t0;
constexpr float c_lv2 = c_cpl - c_cpv;
constexpr float c_ls1 = c_xls + (c_cpi - c_cpv) * c_t0;
constexpr float c_ls2 = c_cpi - c_cpv;
constexpr float c_rp00 = 1.0f / c_p00;
constexpr float c_reps = c_rv / c_rd;
constexpr float c_rddcp = c_rd / c_cp;
constexpr float c_cpdg = c_cp / c_g;
constexpr float c_converge = 0.0002f;
constexpr float c_eps_q = 287.04f / 461.5f;   // getqvs/getqvi local eps (f90:575,592)

template <int TM> inline float getqvs(float p, float t) {      // f90:570-581
  float es = 611.2f * t_exp<TM>(17.67f * (t - 273.15f) / (t - 29.65f));
  return c_eps_q * es / (p - es);
}
template <int TM> inline float getqvi(float p, float t) {      // f90:587-598
  float es = 611.2f * t_exp<TM>(21.8745584f * (t - 273.15f) / (t - 7.66f));
  return c_eps_q * es / (p - es);
}
template <int TM> inline float getthe(float p, float t, float td, float q) {   // f90:604-620
  float tlcl;
  if ((td - t) >= -0.1f) tlcl = t;
  else tlcl = 56.0f + 1.0f / (1.0f / (td - 56.0f) + 0.00125f * t_log<TM>(t / td));
  return t * t_pow<TM>(100000.0f / p, 0.2854f * (1.0f - 0.28f * q)) *
         t_exp<TM>(((3376.0f / tlcl) - 2.54f) * q * (1.0f + 0.81f * q));
}

struct ColOut { float cape, cin, zout; int32_t mulvl; int32_t n_iter, n_sub, status; float cond_b; };

// Conditioning word (SURVEY §8d "ill-conditioned columns"): CIN is only credited when positive area is (re)entered
// (f90:509-523), so at a level where the parcel arrives with b1 < 0 the SIGN of b2 decides between "credit the
// negative area accumulated so far" and "keep accumulating" — the one place where CAPE/CIN are discontinuous in the
// buoyancy.  cond_b = min over such levels of |b2| (m/s2; +inf if there is none): a column whose cond_b is far above
// the arithmetic noise of an implementation cannot flip, one near zero can.  Written to the buffer registered with
// xcape_ref_set_cond_buffer (tests of the tolerance-level `fast` modes classify columns with it).
float* g_cond_out = nullptr;

// Diagnostic trace (profiles/divergence_model.py): per column, a list of int16 — for every layer of the ascent
// -(1000 + k), -nloop, then the number of moist passes of each of its sub-steps.  Used to model what a SIMT warp pays for
// lanes that need different pass counts; not part of any parity check.
thread_local int16_t* g_trace = nullptr;
thread_local int g_trace_cap = 0, g_trace_n = 0;
inline void trace_put(int v) { if (g_trace && g_trace_n < g_trace_cap) g_trace[g_trace_n++] = (int16_t)v; }
enum { ST_OK = 0, ST_SKIPPED = 1, ST_NONCONV = 2 };
constexpr int NLOOP_CAP = 1 << 16;     // sub-steps per layer beyond this, or a NaN step => status 3 (invalid sounding)
constexpr int ITER_BUDGET = 1 << 22;   // moist passes per column beyond this => status 3
enum { ST_INVALID = 3 };

// One column.  pA/tA/tdA hold nk_in levels with stride `ls` (elements).
template <int TM>
void getcape(const float* pA, const float* tA, const float* tdA, int64_t ls_p, int64_t ls,
             float ps_in, float ts_in, float tds_in, float pinc, int source, float ml_depth,
             int adiabat, int nk_in, std::vector<float>& w, ColOut& o) {
  const int nk = nk_in + 1;                                    // f90:219
  w.resize((size_t)7 * (nk + 2));
  float* p = w.data();          // 1-based
  float* t = p + (nk + 2);
  float* td = t + (nk + 2);
  float* pi = td + (nk + 2);
  float* q = pi + (nk + 2);
  float* th = q + (nk + 2);
  float* thv = th + (nk + 2);
  std::vector<float> zv((size_t)nk + 2);
  float* z = zv.data();

  // f90:220-240: prepend surface, convert to mks, pi, q, th, thv
  for (int k = 1; k <= nk; ++k) {
    float pin = (k == 1) ? ps_in : pA[(int64_t)(k - 2) * ls_p];
    float tin = (k == 1) ? ts_in : tA[(int64_t)(k - 2) * ls];
    float tdin = (k == 1) ? tds_in : tdA[(int64_t)(k - 2) * ls];
    p[k] = 100.0f * pin;
    t[k] = 273.15f + tin;
    td[k] = 273.15f + tdin;
    pi[k] = t_pow<TM>(p[k] * c_rp00, c_rddcp);
    q[k] = getqvs<TM>(p[k], td[k]);
    th[k] = t[k] / pi[k];
    thv[k] = th[k] * (1.0f + c_reps * q[k]) / (1.0f + q[k]);
  }
  // f90:244-248
  z[1] = 0.0f;
  for (int k = 2; k <= nk; ++k) {
    float dz = -c_cpdg * 0.5f * (thv[k] + thv[k - 1]) * (pi[k] - pi[k - 1]);
    z[k] = z[k - 1] + dz;
  }
  o.mulvl = -999999;                                            // f90:252-253
  o.zout = -999999.0f;
  o.cape = 0.0f; o.cin = 0.0f;
  o.n_iter = 0; o.n_sub = 0; o.status = ST_OK; o.cond_b = INFINITY;

  int kmax = 1;
  double avgth = 0.0, avgqv = 0.0;
  float th2 = 0, qv2 = 0;
  if (source == 1) {                                            // f90:257-259
    kmax = 1;
  } else if (source == 2) {                                     // f90:261-281
    if (p[1] < 50000.0f) {
      kmax = 1;
    } else {
      float maxthe = 0.0f;
      for (int k = 1; k <= nk; ++k) {
        if (p[k] >= 50000.0f) {
          float the = getthe<TM>(p[k], t[k], td[k], q[k]);
          if (the > maxthe) { o.mulvl = k; maxthe = the; kmax = k; }
        }
      }
    }
  } else {                                                      // f90:284-339 (source 3)
    if ((z[2] - z[1]) > ml_depth) {
      avgth = th[1]; avgqv = q[1]; kmax = 1;
    } else if (z[nk] < ml_depth) {
      avgth = th[nk]; avgqv = q[nk]; kmax = nk;
    } else {
      avgth = 0.0; avgqv = 0.0;
      int k = 2;
      // reference: do while((z(k).le.ml_depth).and.(k.le.nk)) — reads z(nk+1) when
      // z(nk)==ml_depth exactly; guarded here (k bound first).
      while (k <= nk && z[k] <= ml_depth) {
        avgth = avgth + (double)(0.5f * (z[k] - z[k - 1]) * (th[k] + th[k - 1]));
        avgqv = avgqv + (double)(0.5f * (z[k] - z[k - 1]) * (q[k] + q[k - 1]));
        k = k + 1;
      }
      if (k > nk) k = nk;   // only on the guarded edge above
      th2 = th[k - 1] + (th[k] - th[k - 1]) * (ml_depth - z[k - 1]) / (z[k] - z[k - 1]);
      qv2 = q[k - 1] + (q[k] - q[k - 1]) * (ml_depth - z[k - 1]) / (z[k] - z[k - 1]);
      avgth = avgth + (double)(0.5f * (ml_depth - z[k - 1]) * (th2 + th[k - 1]));
      avgqv = avgqv + (double)(0.5f * (ml_depth - z[k - 1]) * (qv2 + q[k - 1]));
      avgth = avgth / (double)ml_depth;
      avgqv = avgqv / (double)ml_depth;
      kmax = 1;
    }
  }

  // f90:355-383 parcel initial state
  float narea = 0.0f;
  int k = kmax;
  float pi2, p2, t2, thv2, b2;
  if (source == 1 || source == 2) {
    th2 = th[kmax]; pi2 = pi[kmax]; p2 = p[kmax]; t2 = t[kmax];
    thv2 = thv[kmax]; qv2 = q[kmax]; b2 = 0.0f;
  } else {
    th2 = (float)avgth; qv2 = (float)avgqv;
    thv2 = th2 * (1.0f + c_reps * qv2) / (1.0f + qv2);
    pi2 = pi[kmax]; p2 = p[kmax]; t2 = th2 * pi2;
    b2 = c_g * (thv2 - thv[kmax]) / thv[kmax];
  }
  float ql2 = 0.0f, qi2 = 0.0f, qt = qv2;
  float cape = 0.0f, cin = 0.0f;
  bool doit = true;
  const bool ice = !(adiabat == 1 || adiabat == 2);
  const bool pseudo = (adiabat == 1 || adiabat == 3);

  // f90:403-559 ascent
  while (doit && k < nk) {
    k = k + 1;
    float b1 = b2;
    float dp = p[k - 1] - p[k];
    int nloop;
    if (dp < pinc) {
      nloop = 1;
    } else {
      float r = dp / pinc;
      // the reference overflows int(dp/pinc) here (undefined behaviour); documented deviation shared with the kernel
      if (!(r < (float)NLOOP_CAP)) { o.cape = 0.0f; o.cin = 0.0f; o.status = ST_INVALID; return; }
      nloop = 1 + (int)r;
      dp = dp / (float)nloop;
    }
    trace_put(-(1000 + k)); trace_put(-nloop);                    // k: index of the layer's upper level in the assembled column
    for (int n = 1; n <= nloop; ++n) {
      float p1 = p2, t1 = t2, th1 = th2, qv1 = qv2, ql1 = ql2, qi1 = qi2;
      p2 = p2 - dp;
      pi2 = t_pow<TM>(p2 * c_rp00, c_rddcp);
      float thlast = th1;
      int i = 0;
      bool not_converged = true;
      o.n_sub++;
      while (not_converged) {
        i = i + 1;
        t2 = thlast * pi2;
        float fliq, fice;
        if (ice) {
          fliq = fmax_(fmin_((t2 - 233.15f) / (273.15f - 233.15f), 1.0f), 0.0f);
          fice = 1.0f - fliq;
        } else { fliq = 1.0f; fice = 0.0f; }
        qv2 = fmin_(qt, fliq * getqvs<TM>(p2, t2) + fice * getqvi<TM>(p2, t2));
        qi2 = fmax_(fice * (qt - qv2), 0.0f);
        ql2 = fmax_(qt - qv2 - qi2, 0.0f);
        float tbar = 0.5f * (t1 + t2);
        float qvbar = 0.5f * (qv1 + qv2);
        float qlbar = 0.5f * (ql1 + ql2);
        float qibar = 0.5f * (qi1 + qi2);
        float lhv = c_lv1 - c_lv2 * tbar;
        float lhs = c_ls1 - c_ls2 * tbar;
        float rm = c_rd + c_rv * qvbar;
        float cpm = c_cp + c_cpv * qvbar + c_cpl * qlbar + c_cpi * qibar;
        th2 = th1 * t_exp_small<TM>(lhv * (ql2 - ql1) / (cpm * tbar) + lhs * (qi2 - qi1) / (cpm * tbar) +
                                    (rm / cpm - c_rddcp) * t_log<TM>(p2 / p1));
        o.n_iter++;
        if (i > 100) {                                           // f90:464-474
          o.cape = 0.0f; o.cin = 0.0f; o.status = ST_NONCONV;
          return;                                                // mulvl/zout keep current values
        }
        if (std::fabs(th2 - thlast) > c_converge) thlast = thlast + 0.3f * (th2 - thlast);
        else not_converged = false;
      }
      trace_put(i);
      if (o.n_iter > ITER_BUDGET) { o.cape = 0.0f; o.cin = 0.0f; o.status = ST_INVALID; return; }
      if (pseudo) { qt = qv2; ql2 = 0.0f; qi2 = 0.0f; }          // f90:487-491
    }
    thv2 = th2 * (1.0f + c_reps * qv2) / (1.0f + qv2 + ql2 + qi2);   // f90:501-503
    b2 = c_g * (thv2 - thv[k]) / thv[k];
    if (b1 < 0.0f) o.cond_b = std::min(o.cond_b, std::fabs(b2));
    float dz = -c_cpdg * 0.5f * (thv[k] + thv[k - 1]) * (pi[k] - pi[k - 1]);
    float parea;
    if (b2 >= 0.0f && b1 < 0.0f) {                               // f90:509-545
      float frac = b2 / (b2 - b1);
      parea = 0.5f * b2 * dz * frac;
      narea = narea - 0.5f * b1 * dz * (1.0f - frac);
      cin = cin + narea;
      narea = 0.0f;
    } else if (b2 < 0.0f && b1 > 0.0f) {
      float frac = b1 / (b1 - b2);
      parea = 0.5f * b1 * dz * frac;
      narea = -0.5f * b2 * dz * (1.0f - frac);
    } else if (b2 < 0.0f) {
      parea = 0.0f;
      narea = narea - 0.5f * dz * (b1 + b2);
    } else {
      parea = 0.5f * dz * (b1 + b2);
      narea = 0.0f;
    }
    cape = cape + fmax_(0.0f, parea);
    if (p[k] <= 10000.0f && b2 < 0.0f) doit = false;             // f90:554-557
    o.zout = z[k];                                               // f90:558
  }
  o.cape = cape; o.cin = cin;
}

template <int TM>
void loopcape_range(int64_t i0, int64_t i1, const float* p3d, const float* t3d, const float* td3d,
                    bool p_is_1d, const float* ps, const float* ts, const float* tds, float pinc,
                    int source, float ml_depth, int adiabat, const int32_t* start_3d, int nk,
                    float* cape, float* cin, int32_t* mulvl, float* zout, int32_t* n_iter,
                    int32_t* n_sub, int32_t* status) {
  std::vector<float> w;
  for (int64_t i = i0; i < i1; ++i) {
    ColOut o;
    if (ts[i] > 0.0f) {                                          // f90:77 / pressure_lev.f90:152
      int ks = start_3d ? start_3d[i] : 1;                       // pressure_lev.f90:154-160
      if (ks < 1) ks = 1;
      if (ks > nk) ks = nk;
      int nk_used = nk - ks + 1;
      const float* pcol = p_is_1d ? p3d + (ks - 1) : p3d + i * nk + (ks - 1);
      getcape<TM>(pcol, t3d + i * nk + (ks - 1), td3d + i * nk + (ks - 1), 1, 1, ps[i], ts[i],
                  tds[i], pinc, source, ml_depth, adiabat, nk_used, w, o);
    } else {
      o.cape = 0; o.cin = 0; o.zout = 0; o.mulvl = 0; o.n_iter = 0; o.n_sub = 0; o.status = ST_SKIPPED; o.cond_b = INFINITY;
    }
    cape[i] = o.cape; cin[i] = o.cin; mulvl[i] = o.mulvl; zout[i] = o.zout;
    if (g_cond_out) g_cond_out[i] = o.cond_b;
    if (n_iter) n_iter[i] = o.n_iter;
    if (n_sub) n_sub[i] = o.n_sub;
    if (status) status[i] = o.status;
  }
}

template <class F> void par_for(int64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 2) { f(0, n); return; }
  std::vector<std::thread> th;
  int64_t blk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    int64_t a = std::min<int64_t>(n, t * blk), b = std::min<int64_t>(n, a + blk);
    if (a < b) th.emplace_back([=] { f(a, b); });
  }
  for (auto& x : th) x.join();
}

// ---------------------------------------------------------------------------
// stdheight  (stdheight_2D_model_lev.f90:74-160; pressure: stdheight_2D_pressure_lev.f90:65-93,133-219)
// binary64 arithmetic, single-precision literals promoted
// ---------------------------------------------------------------------------
constexpr double h_R = (double)287.04f, h_g = (double)-9.80665f, h_eps = (double)0.6219800858985514f;
constexpr double h_t0 = (double)273.15f, h_c1 = (double)6.112f, h_c2 = (double)53.49f, h_c3 = (double)5.09f;

inline double tvirt(double T, double Td, double P) {
  double Tin = T + h_t0;
  double Tdin = Td + h_t0;
  double E = h_c1 * std::exp((h_c2 - (6808 / Tdin) - h_c3 * std::log(Tdin)));
  double w = h_eps * (E / (P - E));
  return Tin * ((w + h_eps) / (h_eps * (1 + w)));
}
// P,T,Td: nk levels, unit stride; H out nk
void stdheight_col(const double* P, const double* T, const double* Td, double Ps, double Ts,
                   double Tds, double Hin, int nk, double* H, double* Hs) {
  double Tvs = tvirt(Ts, Tds, Ps);
  *Hs = Hin;
  double Tvprev = Tvs, Hprev = Hin, Pprev = Ps;
  for (int k = 0; k < nk; ++k) {
    double Tv = tvirt(T[k], Td[k], P[k]);
    double h = Hprev + ((h_R * ((Tv + Tvprev) / 2) / h_g)) * (std::log(P[k] / Pprev));
    H[k] = h; Hprev = h; Tvprev = Tv; Pprev = P[k];
  }
}

// ---------------------------------------------------------------------------
// Bunkers (Bunkers_model_lev.f90:75-180, DINTERP2DZ :188-236) — binary32
// ---------------------------------------------------------------------------
float dinterp2dz(const float* V, const float* Z, float height, int nz) {   // 0-based arrays of nz
  float out = -999999.0f;
  int ip = 0, im = 1;
  if (Z[0] > Z[nz - 1]) { ip = 1; im = 0; }
  for (int kp = nz; kp >= 2; --kp) {                     // KP is 1-based in the reference
    float zlo = Z[kp - im - 1], zhi = Z[kp - ip - 1];
    if (zlo <= height && zhi > height) {
      float w2 = (height - zlo) / (zhi - zlo);
      float w1 = (float)(1.0 - (double)w2);              // 1.D0 - W2 (f90:228)
      out = w1 * V[kp - im - 1] + w2 * V[kp - ip - 1];
      break;
    }
  }
  return out;
}
void bunkers_col(const float* U, const float* V, const float* Z, int nk, float* RM, float* LM, float* M6) {
  float us[13], vs[13];
  us[0] = U[0]; vs[0] = V[0];
  for (int i = 1; i < 13; ++i) {
    float lvl = 500.0f * (float)i;
    us[i] = dinterp2dz(U, Z, lvl, nk);
    vs[i] = dinterp2dz(V, Z, lvl, nk);
  }
  float mu = 0.0f, mv = 0.0f;                            // f2py zero-fills intent(out) (SURVEY B-8)
  for (int i = 0; i < 13; ++i) { mu = mu + us[i]; mv = mv + vs[i]; }
  mu = mu / 13.0f; mv = mv / 13.0f;
  float uu = 0.0f, vu = 0.0f, ud = 0.0f, vd = 0.0f;      // uninitialised in the reference; 0 reproduces goldens
  for (int i = 11; i < 13; ++i) { uu = uu + us[i]; vu = vu + vs[i]; }
  uu = uu / 2.0f; vu = vu / 2.0f;
  for (int i = 0; i < 2; ++i) { ud = ud + us[i]; vd = vd + vs[i]; }
  ud = ud / 2.0f; vd = vd / 2.0f;
  float ushr = uu - ud, vshr = vu - vd;
  float nrm = sqrtf(ushr * ushr + vshr * vshr);          // (..)**0.5
  RM[0] = mu + 7.5f * vshr / nrm;
  RM[1] = mv - 7.5f * ushr / nrm;
  LM[0] = mu - 7.5f * vshr / nrm;
  LM[1] = mv + 7.5f * ushr / nrm;
  M6[0] = mu; M6[1] = mv;
}

// ---------------------------------------------------------------------------
// SREH (SREH_model_lev.f90:61-129) — binary64
// ---------------------------------------------------------------------------
inline double interp1(double y1, double y3, double x1, double x2, double x3) {
  if (x3 == x1) x1 = x1 - (double)0.01f;
  return y1 + ((y3 - y1) * ((x2 - x1) / (x3 - x1)));
}
void sreh_col(const double* u, const double* v, const double* z, double curm, double cvrm,
              double culm, double cvlm, double top, int nk, double* srm, double* slm,
              std::vector<double>& w) {
  w.resize((size_t)2 * nk);
  double* ut = w.data(); double* vt = ut + nk;
  for (int k = 0; k < nk; ++k) { ut[k] = u[k]; vt[k] = v[k]; }
  int ktop = 0;                                           // 1-based
  for (int k = 2; k <= nk; ++k) {
    if (z[k - 1] > top && ktop == 0) {
      ktop = k;
      ut[k - 1] = interp1(u[k - 1], u[k - 2], z[k - 1], top, z[k - 2]);
      vt[k - 1] = interp1(v[k - 1], v[k - 2], z[k - 1], top, z[k - 2]);
    }
  }
  double s = 0.0;
  for (int k = 2; k <= ktop; ++k)
    s = s + (((ut[k - 1] - curm) * (vt[k - 1] - vt[k - 2])) - ((vt[k - 1] - cvrm) * (ut[k - 1] - ut[k - 2])));
  *srm = -s;
  s = 0.0;
  for (int k = 2; k <= ktop; ++k)
    s = s + (((ut[k - 1] - culm) * (vt[k - 1] - vt[k - 2])) - ((vt[k - 1] - cvlm) * (ut[k - 1] - ut[k - 2])));
  *slm = -s;
}

}  // namespace

// ===========================================================================
// C entry points.  Array layout = the f2py layout: (nk, n2) Fortran order, i.e.
// element (k, i) at [i*nk + k]  (each column contiguous).
// ===========================================================================
extern "C" {

// CAPE_CODE_model_lev.pyf:6-24 / f90:4-91
int xcape_ref_loopcape_ml(const float* p3d, const float* t3d, const float* td3d, const float* ps,
                          const float* ts, const float* tds, float pinc, int source, float ml_depth,
                          int adiabat, int nk, int64_t n2, float* cape, float* cin, int32_t* mulvl,
                          float* zout, int tmode, int nthreads, int32_t* n_iter, int32_t* n_sub,
                          int32_t* status) {
  if (source < 1 || source > 3 || adiabat < 1 || adiabat > 4 || !(pinc > 0.0f) || nk < 1) return 1;
  par_for(n2, nthreads, [=](int64_t a, int64_t b) {
    if (tmode == T_LIBM) loopcape_range<T_LIBM>(a, b, p3d, t3d, td3d, false, ps, ts, tds, pinc, source, ml_depth, adiabat, nullptr, nk, cape, cin, mulvl, zout, n_iter, n_sub, status);
    else if (tmode == T_CR) loopcape_range<T_CR>(a, b, p3d, t3d, td3d, false, ps, ts, tds, pinc, source, ml_depth, adiabat, nullptr, nk, cape, cin, mulvl, zout, n_iter, n_sub, status);
    else loopcape_range<T_SPEC>(a, b, p3d, t3d, td3d, false, ps, ts, tds, pinc, source, ml_depth, adiabat, nullptr, nk, cape, cin, mulvl, zout, n_iter, n_sub, status);
  });
  return 0;
}

// CAPE_CODE_pressure_lev.pyf:26-45 / f90:88-169   (p is (nk,1); start_3d 1-based)
int xcape_ref_loopcape_pl1d(const float* t3d, const float* td3d, const float* p1d, const float* ps,
                            const float* ts, const float* tds, float pinc, int source,
                            float ml_depth, int adiabat, const int32_t* start_3d, int nk,
                            int64_t n2, float* cape, float* cin, int32_t* mulvl, float* zout,
                            int tmode, int nthreads, int32_t* n_iter, int32_t* n_sub,
                            int32_t* status) {
  if (source < 1 || source > 3 || adiabat < 1 || adiabat > 4 || !(pinc > 0.0f) || nk < 1) return 1;
  par_for(n2, nthreads, [=](int64_t a, int64_t b) {
    if (tmode == T_LIBM) loopcape_range<T_LIBM>(a, b, p1d, t3d, td3d, true, ps, ts, tds, pinc, source, ml_depth, adiabat, start_3d, nk, cape, cin, mulvl, zout, n_iter, n_sub, status);
    else if (tmode == T_CR) loopcape_range<T_CR>(a, b, p1d, t3d, td3d, true, ps, ts, tds, pinc, source, ml_depth, adiabat, start_3d, nk, cape, cin, mulvl, zout, n_iter, n_sub, status);
    else loopcape_range<T_SPEC>(a, b, p1d, t3d, td3d, true, ps, ts, tds, pinc, source, ml_depth, adiabat, start_3d, nk, cape, cin, mulvl, zout, n_iter, n_sub, status);
  });
  return 0;
}

// register (or clear, with nullptr) the per-column output buffer of the conditioning word for the NEXT loopcape call
void xcape_ref_set_cond_buffer(float* buf) { g_cond_out = buf; }

// diagnostic: loopcape_pl1d / loopcape_ml (SPEC arithmetic, one thread) with the per-sub-step pass counts of every
// column written to trace[i*cap .. ) (see g_trace above; unused slots stay 0)
int xcape_ref_cape_trace(const float* p, const float* t3d, const float* td3d, int p_is_1d, const float* ps,
                         const float* ts, const float* tds, float pinc, int source, float ml_depth, int adiabat,
                         const int32_t* start_3d, int nk, int64_t n2, int16_t* trace, int cap) {
  if (source < 1 || source > 3 || adiabat < 1 || adiabat > 4 || !(pinc > 0.0f) || nk < 1) return 1;
  std::vector<float> cape(1), cin(1), zout(1);
  std::vector<int32_t> mulvl(1);
  std::memset(trace, 0, (size_t)n2 * cap * sizeof(int16_t));
  for (int64_t i = 0; i < n2; ++i) {
    g_trace = trace + i * cap; g_trace_cap = cap; g_trace_n = 0;
    loopcape_range<T_SPEC>(i, i + 1, p, t3d, td3d, p_is_1d != 0, ps, ts, tds, pinc, source, ml_depth, adiabat, start_3d, nk,
                           cape.data() - i, cin.data() - i, mulvl.data() - i, zout.data() - i, nullptr, nullptr, nullptr);
  }
  g_trace = nullptr;
  return 0;
}

// stdheight_2D_model_lev.pyf:6-19 / f90:4-34
int xcape_ref_loop_stdheight_ml(const double* P, const double* T, const double* Td, const double* Ps,
                                const double* Ts, const double* Tds, const double* Hin, int nk,
                                int64_t nx, double* H, double* Hs, int nthreads) {
  par_for(nx, nthreads, [=](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i)
      stdheight_col(P + i * nk, T + i * nk, Td + i * nk, Ps[i], Ts[i], Tds[i], Hin[i], nk, H + i * nk, Hs + i);
  });
  return 0;
}

// stdheight_2D_pressure_lev.pyf:21-35 / f90:65-93  (start_3d arrives as double)
int xcape_ref_loop_stdheight_pl1d(const double* T, const double* Td, const double* P1d,
                                  const double* Ps, const double* Ts, const double* Tds,
                                  const double* Hin, const double* start_3d, int nk, int64_t nx,
                                  double* H, double* Hs, int nthreads) {
  par_for(nx, nthreads, [=](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      int ks = (int)start_3d[i];
      if (ks < 1) ks = 1;
      if (ks > nk) ks = nk;
      for (int k = 0; k < ks - 1; ++k) H[i * nk + k] = -999999;
      stdheight_col(P1d + (ks - 1), T + i * nk + (ks - 1), Td + i * nk + (ks - 1), Ps[i], Ts[i],
                    Tds[i], Hin[i], nk - ks + 1, H + i * nk + (ks - 1), Hs + i);
    }
  });
  return 0;
}

// Bunkers_model_lev.pyf:8-21 (start_3d == NULL) / Bunkers_pressure_lev.pyf:6-20 (start_3d real)
// RM/LM/Mean6 are (2, n2) Fortran order: element (c, i) at [i*2 + c].
int xcape_ref_bunkers_loop(const float* U, const float* V, const float* Z, const float* Us,
                           const float* Vs, const float* Zs, const float* start_3d, int nk,
                           int64_t n2, float* RM, float* LM, float* M6, int nthreads) {
  par_for(n2, nthreads, [=](int64_t a, int64_t b) {
    std::vector<float> ua(nk + 1), va(nk + 1), za(nk + 1);
    for (int64_t i = a; i < b; ++i) {
      int ks = start_3d ? (int)start_3d[i] : 1;
      if (ks < 1) ks = 1;
      if (ks > nk) ks = nk;
      int n = nk - ks + 1;
      ua[0] = Us[i]; va[0] = Vs[i]; za[0] = Zs[i];
      for (int k = 0; k < n; ++k) { ua[k + 1] = U[i * nk + ks - 1 + k]; va[k + 1] = V[i * nk + ks - 1 + k]; za[k + 1] = Z[i * nk + ks - 1 + k]; }
      bunkers_col(ua.data(), va.data(), za.data(), n + 1, RM + 2 * i, LM + 2 * i, M6 + 2 * i);
    }
  });
  return 0;
}

// SREH_model_lev.pyf:6-23 (start_3d == NULL) / SREH_pressure_lev.pyf:6-24 (start_3d double)
int xcape_ref_loop_sreh(const double* u, const double* v, const double* z, const double* us,
                        const double* vs, const double* zs, const double* cu_rm, const double* cv_rm,
                        const double* cu_lm, const double* cv_lm, double top, const double* start_3d,
                        int nk, int64_t n2, double* srm, double* slm, int nthreads) {
  par_for(n2, nthreads, [=](int64_t a, int64_t b) {
    std::vector<double> ua(nk + 1), va(nk + 1), za(nk + 1), w;
    for (int64_t i = a; i < b; ++i) {
      int ks = start_3d ? (int)start_3d[i] : 1;
      if (ks < 1) ks = 1;
      if (ks > nk) ks = nk;
      int n = nk - ks + 1;
      ua[0] = us[i]; va[0] = vs[i]; za[0] = zs[i];
      for (int k = 0; k < n; ++k) { ua[k + 1] = u[i * nk + ks - 1 + k]; va[k + 1] = v[i * nk + ks - 1 + k]; za[k + 1] = z[i * nk + ks - 1 + k]; }
      sreh_col(ua.data(), va.data(), za.data(), cu_rm[i], cv_rm[i], cu_lm[i], cv_lm[i], top, n + 1, srm + i, slm + i, w);
    }
  });
  return 0;
}

// scalar probes of the transcendental modes (tests compare SPEC vs CR vs LIBM)
float xcape_ref_expf_small(float x, int tmode) { return tmode == 0 ? t_exp_small<0>(x) : tmode == 1 ? t_exp_small<1>(x) : t_exp_small<2>(x); }
void xcape_ref_expf_small_v(const float* x, float* y, int64_t n, int tmode) { for (int64_t i = 0; i < n; ++i) y[i] = xcape_ref_expf_small(x[i], tmode); }
float xcape_ref_expf(float x, int tmode) { return tmode == 0 ? t_exp<0>(x) : tmode == 1 ? t_exp<1>(x) : t_exp<2>(x); }
float xcape_ref_logf(float x, int tmode) { return tmode == 0 ? t_log<0>(x) : tmode == 1 ? t_log<1>(x) : t_log<2>(x); }
float xcape_ref_powf(float x, float y, int tmode) { return tmode == 0 ? t_pow<0>(x, y) : tmode == 1 ? t_pow<1>(x, y) : t_pow<2>(x, y); }
void xcape_ref_expf_v(const float* x, float* y, int64_t n, int tmode) { for (int64_t i = 0; i < n; ++i) y[i] = xcape_ref_expf(x[i], tmode); }
void xcape_ref_logf_v(const float* x, float* y, int64_t n, int tmode) { for (int64_t i = 0; i < n; ++i) y[i] = xcape_ref_logf(x[i], tmode); }
void xcape_ref_powf_v(const float* x, const float* e, float* y, int64_t n, int tmode) { for (int64_t i = 0; i < n; ++i) y[i] = xcape_ref_powf(x[i], e[i], tmode); }

int xcape_ref_fp_contract(void) {
#ifdef XC_ORACLE_CONTRACT
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"
