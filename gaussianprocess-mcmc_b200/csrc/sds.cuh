// Per-wave state of the device-resident surrogate-data slice-sampling sweep (see sds.cu).
#pragma once
#include "common.cuh"

namespace gpmc {

struct SdsState {
    int n, P, ldv;                 // observations, hyper-parameters, row stride of the per-chain vectors
    int iter;                      // MCMC iteration (burn-in switch at 500, sliceSample.py:128,133,151)
    unsigned sweep;                // RNG stream counter (usually == iter)
    unsigned chain0;               // global id of the wave's first chain (RNG key)
    unsigned long long seed;
    double my, lower, upper;       // mean(y), 0 - my, 100 - my   (:102,114-115)
    const double *y;               // [n]
    const double *scale;           // [P]
    const double *prior_k, *prior_theta;   // [P]
    // chain state (wave-local, row c)
    const double *F;               // [B, ldv] current latent f
    const double *hyp;             // [B, P]   current hyper-parameters
    double *F_out;                 // [B, ldv] accepted f'
    double *hyp_out;               // [B, P]
    double *loglik_out;            // [B]      log N(g; 0, K+S) at the accepted theta
    // work vectors
    double *g, *svec, *fprop;      // [B, ldv]
    double *theta;                 // [B, P]   hyper-parameters the aux model is evaluated at
    double *hyp_min, *hyp_max;     // [B, P]
    double *G;                     // [B]      log marginal of the last evaluation (curG / propG)
    double *log_u0, *threshold, *cur_llk, *curG, *last_proposal, *last_llk;   // [B]
    int *done, *ntrips;            // [B]
    int *map, *count;              // active list
    // ---- resident loop only (gpmc_sds_sweep, default path; all nullptr / 0 on the wave path).  Rows of the work
    // arrays above are SLOTS; a slot holds chain chain_of[slot] of the call (or -1), chains are admitted into free
    // slots on the device, and inputs / results move between the caller's arrays and the slot rows inside the kernels.
    int *chain_of;                 // [cap] chain of the call held by the slot, -1 = free
    int *phase;                    // [cap] SDS_PHASE_*
    int *parked, *resolved;        // [cap] factorisation failed: wait for the host-side jitter ladder / ladder done
    const int *info1, *info2;      // [cap] status of chol(K+S) and chol(R + 1e-11 I) of the evaluation just queued
    int *map_new, *count_new;      // slots admitted this round (evaluate at the current theta)
    int *map_act, *count_act;      // slots in their shrink loop (evaluate at a proposal)
    int *status_word;              // [8]: count, count_new, count_act, parked, next_chain, round
    int *next_chain;               // next chain of the call to admit
    int n_chains, cap, max_trips;
    const double *F_glob;          // [n_chains, n] caller's latent vectors (in; overwritten on accept)
    double *F_glob_out;
    const double *hyp_glob;        // [n_chains, P]
    double *hyp_glob_out;
    int *ntrips_glob, *status_glob;   // [n_chains] (may be nullptr)
    double *loglik_glob;           // [n_chains] (may be nullptr)
    double *hyp_stage;             // [cap, P] slot copy of the chain's current theta (== hyp on this path)
    double *F_stage;               // [cap, ldv] slot copy of the chain's current f (== F on this path)
    // ---- many iterations per call (gpmc_sds_run): a slot keeps its chain from iteration iter to iter + n_iters - 1, every
    // chain advances on its own (no chain waits for the slowest one at an iteration boundary), history goes to hist_*.
    int *iter_of;                  // [cap] MCMC iteration the slot's chain is at (nullptr: one transition per call, `iter`)
    int n_iters;                   // iterations per chain in this call
    double *hist_hyp;              // [n_chains, n_iters, P]   (may be nullptr)
    double *hist_loglik;           // [n_chains, n_iters]      (may be nullptr)
    int *hist_trips;               // [n_chains, n_iters]      (may be nullptr)
    double *hist_f;                // [n_chains, n_keep, n], f after every thin-th iteration (may be nullptr)
    int thin, n_keep;
    int *n_exhausted;              // transitions that used up max_trips (state kept)
    // explicit randomness (all nullptr -> Philox)
    const double *tape_z;          // [B, n]
    const double *tape_v;          // [B, P]
    const double *tape_u0;         // [B]
    const double *tape_U;          // [B, tape_trips, P]
    int tape_trips;
};

enum { SDS_PHASE_FREE = 0, SDS_PHASE_NEW = 1, SDS_PHASE_ACTIVE = 2 };
enum { SDS_SW_COUNT = 0, SDS_SW_NEW = 1, SDS_SW_ACT = 2, SDS_SW_PARKED = 3, SDS_SW_NEXT = 4, SDS_SW_ROUND = 5, SDS_SW_WORDS = 8 };

int launch_sds_begin(const SdsState &st, int nchains, cudaStream_t s);
int launch_sds_threshold(const SdsState &st, int nchains, cudaStream_t s);
int launch_sds_propose(const SdsState &st, int nactive, int trip, cudaStream_t s);
int launch_sds_accept(const SdsState &st, int nactive, cudaStream_t s);
int launch_sds_compact(const SdsState &st, int nchains, cudaStream_t s);
// resident loop: admit waiting chains into free slots, rebuild the three slot lists, publish the status word
int launch_sds_admit(const SdsState &st, int round, cudaStream_t s);

}  // namespace gpmc
