// fwd_kernels.cuh -- launch interface of the fused forward (transmit) path (fwd_kernels.cu)
#pragma once
#include "spec.cuh"

namespace amps {

constexpr int kFwdMaxCar   = 3;
constexpr int kFwdTileSym  = 63;                 // new 100 kS/s symbols per tile -> 6300 output samples
constexpr int kFwdMaxTap4  = 81;                 // taps per polyphase arm (321 padded to 324)
constexpr int kFwdHistLen  = kFwdMaxTap4 + 1;    // symbols of history a call needs from the previous one
constexpr int kFwdThreads  = 256;
constexpr int kFwdScanBlock = 4096;              // symbols per block of the prefix-sum kernels
constexpr int kFwdInterp   = 100;                // 10 MS/s / 100 kS/s = 4 (reference) x 25 (CIC)
constexpr int kFwdVoiceLegs = 2;                 // voice legs: one gated by audio_mute (+60 kHz), one always on (+90 kHz)
constexpr int kFwdVoicePer  = 32;                // max taps per arm of the x25 voice resampler (225 taps / 8 arms -> 29)
constexpr int kFwdVoiceHist = 32;                // phasors of history a call needs from the previous one

struct FwdScanParams {
    const uint8_t *sym[kFwdMaxCar];
    int32_t       *sloc[kFwdMaxCar];             // inclusive scan inside each 4096-symbol block
    int32_t       *btot[kFwdMaxCar];             // block totals
    int32_t       *boff[kFwdMaxCar];             // exclusive scan of the block totals + carry
    int32_t       *carry;                        // [kFwdMaxCar] running sum across calls
    const uint8_t *hsym_old[kFwdMaxCar];
    const int32_t *hS_old[kFwdMaxCar];
    uint8_t       *hsym_new[kFwdMaxCar];
    int32_t       *hS_new[kFwdMaxCar];
    uint32_t       nsym;
};

struct FwdParams {
    const uint8_t *sym[kFwdMaxCar];
    const int32_t *sloc[kFwdMaxCar];
    const int32_t *boff[kFwdMaxCar];
    const uint8_t *hsym[kFwdMaxCar];             // previous call's last kFwdHistLen symbols
    const int32_t *hS[kFwdMaxCar];               // and their absolute phase sums
    float2        *out;
    uint32_t       nsym;
    int            ncar;
    uint32_t       fcw_fm;                       // FM phase step per unit symbol (2^32 = one turn)
    uint32_t       m_base;                       // absolute 400 kS/s index of this call's first sample (mod 2^32)
    uint32_t       fcw_mix25[kFwdMaxCar];        // mixer phase step per 400 kS/s sample (25 * fcw)
    float          scale;
    int            ntap4[kFwdMaxCar];
    float2         w25[kFwdMaxCar];              // mixer phasor of one 400 kS/s step
    float2         C1[kFwdMaxCar][15];           // 5 cic5[u] e^{j phi_c(5 u)}: first x5 CIC^3 stage carrying the mixer
    float2         C1j[kFwdMaxCar][15];          // j C1 = (-C1.im, C1.re), ready-made: a uniform-register FFMA2 operand instead of a
                                                 // negate + move per tap
    float          G2[15];                       // out_scale * 5 cic5[u]: second (shared) x5 CIC^3 stage
    float          taps[kFwdMaxCar][4 * kFwdMaxTap4];
    // polyphase work split of fwd_fused_kernel<false>: warp w runs taps [k0, k1) of carrier c for all 64 symbols of the
    // tile; slot 0 of a carrier owns the result, slots 1.. hand their partial sums over through shared memory and are
    // added in slot order (a fixed order: a sample's value does not depend on tiling or chunking)
    struct Seg { int8_t c, slot, nslot, hidx; int16_t k0, k1; } seg[kFwdThreads / 32];   // hidx: helper slot in smem (owner: its first helper's)
    // ---- voice legs (fwd_fused_kernel<true> only): nbfm_tx phasors @16 kS/s, x25 arb resampler, added to a carrier's
    //      400 kS/s samples before its mixer (grc/ampsbs.grc:4494-4500, 4632-4638)
    const float2  *vph[kFwdVoiceLegs];           // this call's phasors (4 nsym / 25 of them), already muted where gated
    const float2  *vhist[kFwdVoiceLegs];         // previous call's last kFwdVoiceHist phasors
    int            vcar[kFwdVoiceLegs];          // carrier whose mixer the leg shares, -1 = leg unused
    int            vper;                         // taps per arm of the x25 resampler (<= kFwdVoicePer)
    uint32_t       n_audio;                      // audio samples of this call
    const float   *Eg;                           // device copy of the resampler taps, E[r * kFwdVoicePer + k]
};

// ---- voice pre-pass: audio (+SAT) -> pre-emphasis -> FM phase -> phasors @16 kS/s
constexpr int kVoiceImp = 192;                   // terms of the pre-emphasis impulse response kept (|pole|^192 ~ 1e-20)
struct VoicePrepParams {
    const float *audio;                          // this call's samples
    const double *hx_old;                        // previous call's last kVoiceImp inputs (audio + SAT), zeros at stream start
    double      *hx_new;
    unsigned long long *delta;                   // FM phase steps (2^64 = one turn), inclusive-scanned inside each 256-sample block
    unsigned long long *btot, *boff;             // block totals / absolute phase in front of each block
    unsigned long long *phase;                   // running FM phase (device scalar, carried between calls)
    float2      *vph[kFwdVoiceLegs];
    const float2 *vhist_old[kFwdVoiceLegs];
    float2      *vhist_new[kFwdVoiceLegs];
    int          leg_muted[kFwdVoiceLegs];       // 1 = this call's phasors of the leg are zeros (mute_xx before the resampler)
    unsigned long long a_base;                   // absolute index of audio[0] (selects the SAT phase)
    uint32_t     n_audio;
    double       cycles_per_unit;                // max_dev / audio_rate
    double       g[kVoiceImp];
    double       sat[8];                         // sat_amp * cos(2 pi 3 k / 8): 6 kHz at 16 kS/s
};
cudaError_t launch_voice_prep(const VoicePrepParams &p, cudaStream_t st);
cudaError_t launch_fwd_fused_voice(const FwdParams &p, int grid, cudaStream_t st);

// ---- Manchester-bit fast path: input is one byte per 10 kbit/s data bit (0, 1, 0xFF = muted) instead of half-symbol samples
constexpr int kFbTileBits  = 5;                  // bits per tile -> 200 samples @400 kS/s -> 5000 output samples
constexpr int kFbSymPerBit = 10;                 // 100 kS/s FM samples per bit (2 half-symbols x 5)
constexpr int kFbMPerBit   = 40;                 // 400 kS/s samples per bit
constexpr int kFbOutPerBit = 1000;               // 10 MS/s samples per bit
constexpr int kFbRespBits  = 9;                  // bits whose response overlaps one 400 kS/s sample: ceil((40 + 320) / 40)
constexpr int kFbRespLen   = kFbRespBits * kFbMPerBit;   // 360
constexpr int kFbHistBits  = 9;                  // bits of history a call needs from the previous one
constexpr int kFbTileM     = kFbTileBits * kFbMPerBit;   // 200
// Grouped form of the per-bit response, used while none of a tile's bits is muted.  A Manchester 1 is the mirror image of
// a 0, so its FM waveform -- and, the interpolator taps being real, its response -- is the complex conjugate: R1 = conj(R0).
// The real part of a 400 kS/s sample therefore does not depend on the data at all, and the imaginary part is a signed sum
// that is looked up three bits at a time.  Per carrier: RW[40] = (sum_d Re R0[u+40d]) * w40[u] (float2), JW[40] = j w40[u]
// (float2), I3[3][8][40] = sum_{k<3} (+-) Im R0[u + 40 (8 - 3G - k)] (float), sign - where bit k of the pattern is 1.
constexpr int kFbFastLen   = 2 * 2 * kFbMPerBit + 3 * 8 * kFbMPerBit;   // 1120 floats

struct FwdBitsParams {
    const uint8_t *bits[kFwdMaxCar];             // this call's bits
    const uint8_t *hbits[kFwdMaxCar];            // previous call's last kFbHistBits bits (0xFF at stream start)
    const float2  *resp;                         // [ncar][2][kFbRespLen] response of the x4 interpolator to one Manchester bit
                                                 // (global memory; only tiles that straddle a mute transition walk it)
    const float   *fast;                         // [ncar][kFbFastLen] grouped tables of the same response (see kFbFastLen)
    float2        *out;
    uint32_t       nbits;
    int            ncar;
    unsigned long long bit_base;                 // absolute index of this call's first bit
    uint32_t       fcw_mix1000[kFwdMaxCar];      // mixer phase step per bit (1000 output samples)
    float2         w40[kFwdMaxCar][kFbMPerBit];  // mixer phasors of the 40 400-kS/s steps inside a bit
    float2         C1[kFwdMaxCar][15];
    float          G2[15];
};

size_t fwd_bits_smem_bytes();
cudaError_t launch_fwd_bits(const FwdBitsParams &p, int grid, cudaStream_t st);
size_t fwd_smem_bytes();
cudaError_t fwd_configure_device();
cudaError_t launch_fwd_scan(const FwdScanParams &p, int ncar, cudaStream_t st);
cudaError_t launch_fwd_fused(const FwdParams &p, int grid, cudaStream_t st);

}  // namespace amps
