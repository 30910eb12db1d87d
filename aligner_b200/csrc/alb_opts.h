// alb_opts.h -- tuning / test options of libaligner_b200.so.
//
// The options are read from the environment ONCE, when the library is first used, and can be changed afterwards with
// alb200_set_option() (include/aligner_b200.h); no launch path calls getenv.  `gen` changes with every update so that the
// per-thread configuration caches can tell.
#pragma once
namespace alb {
struct Opts {
    char force[48];        // "R,TF,NS,bits_smem,skew,cluster": pins the MAS kernel shape (tests, sweeps); "" = heuristics
    int latency_max_b;     // >= 0: batch size up to which the one-CTA-per-SM (latency) configuration is taken
    int tmap_promo;        // -1 default (128 B), 0 none, 1 64 B, 2 128 B, 3 256 B: L2 promotion of the score tensor maps
    int no_tail_box;       // 1: no short TMA box for the partly filled last warp
    int force_unaligned;   // 1: take the element-wise loader even for 16-byte aligned rows
    int dbg;               // 1: per-warp clock64 stamps on stderr (needs a -DALB200_DBG_BUILD=1 library)
    int nc_ffma;           // 1: CUDA-core score kernels (cross-check of the tensor-core path)
    int nc_v1;             // 1: first-generation tcgen05 score kernels (operands fetched with plain loads)
    int fused_seq;         // fused score + search entry: 1 = always back to back, 2 = pipelined whenever possible (default: by batch size)
    int nc_no_pdl;         // 1: the score kernel waits for the whole prep kernel (no programmatic dependent launch)
    int no_shared_zero;    // 1: every search CTA zero-fills its own utterance's output (no filler CTAs on the idle SMs)
    unsigned gen;
};
const Opts& opts();
}  // namespace alb
