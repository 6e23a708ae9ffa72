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
    // explicit randomness (all nullptr -> Philox)
    const double *tape_z;          // [B, n]
    const double *tape_v;          // [B, P]
    const double *tape_u0;         // [B]
    const double *tape_U;          // [B, tape_trips, P]
    int tape_trips;
};

int launch_sds_begin(const SdsState &st, int nchains, cudaStream_t s);
int launch_sds_threshold(const SdsState &st, int nchains, cudaStream_t s);
int launch_sds_propose(const SdsState &st, int nactive, int trip, cudaStream_t s);
int launch_sds_accept(const SdsState &st, int nactive, cudaStream_t s);
int launch_sds_compact(const SdsState &st, int nchains, cudaStream_t s);

}  // namespace gpmc
