/* A plain C99 client of include/conclave_b200.h — no Python, no torch: what a maintainer binding the library from C
 * (or through cgo / JNI / any FFI) would write.  Uploads a synthetic roster, runs one batch of trials on the engine
 * given on the command line and prints "winner rounds" per trial followed by the histograms.
 * With a 7th argument (a device count) the same batch goes through csim_group_*: one process, that many GPUs, trials sharded
 * by sim id, histograms summed by NCCL — and the output must be identical to the single-context run.
 * usage: c_client <n_electors> <n_trials> <engine 0..3> <wiring 0|1> <precision 0|1> <seed> [n_devices] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "conclave_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        csim_status st_ = (call);                                                    \
        if (st_ != CSIM_OK) {                                                        \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, csim_last_error());   \
            return 2;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    if (argc != 7 && argc != 8) { fprintf(stderr, "usage: %s E n engine wiring precision seed [n_devices]\n", argv[0]); return 1; }
    const int E = atoi(argv[1]);
    const uint64_t n = strtoull(argv[2], NULL, 10);
    const int engine = atoi(argv[3]), wiring = atoi(argv[4]), precision = atoi(argv[5]);
    const uint64_t seed = strtoull(argv[6], NULL, 10);
    const int n_devices = argc == 8 ? atoi(argv[7]) : 0;          /* 0: one context on device 0, no group */

    /* deterministic synthetic roster (a tiny LCG: the Python side of the test regenerates the same numbers) */
    double* x = malloc(sizeof(double) * E);
    int32_t* region = malloc(sizeof(int32_t) * E);
    uint8_t* pap = malloc(E);
    uint32_t s = 12345u;
    for (int i = 0; i < E; ++i) {
        s = s * 1664525u + 1013904223u; x[i] = (double)(s >> 8) / 16777216.0;
        s = s * 1664525u + 1013904223u; region[i] = (int32_t)((s >> 16) % 8u);
        s = s * 1664525u + 1013904223u; pap[i] = (uint8_t)(((s >> 16) % 100u) < 15u);
    }

    csim_ctx* ctx = NULL;
    csim_group* grp = NULL;
    if (n_devices > 0) {
        int32_t devs[64], visible = 0;
        CHECK(csim_device_count(&visible));
        if (n_devices > 64 || n_devices > visible) {
            fprintf(stderr, "%d devices asked, %d visible\n", n_devices, (int)visible);
            return 3;
        }
        for (int i = 0; i < n_devices; ++i) devs[i] = i;
        CHECK(csim_group_create(&grp, devs, n_devices));
        CHECK(csim_group_upload_roster(grp, E, x, region, pap));
        if (csim_group_size(grp) != n_devices) { fprintf(stderr, "group size %d\n", (int)csim_group_size(grp)); return 3; }
    } else {
        CHECK(csim_create(&ctx, 0));
        CHECK(csim_upload_roster(ctx, E, x, region, pap));
    }

    csim_params p;
    memset(&p, 0, sizeof p);                       /* README parameters (BASELINE config 2) */
    p.initial_beta_weight = 1.5; p.enable_dynamic_beta = 1; p.beta_mode = CSIM_BETA_STEPPED;
    p.beta_increment_amount = 0.3; p.beta_increment_interval_rounds = 1; p.beta_growth_rate = 1.05;
    p.stickiness_factor = 0.8; p.bandwagon_strength = 0.7; p.regional_affinity_bonus = 0.1; p.papabile_weight_factor = 2.0;
    p.fatigue_vote_share_threshold = 0.05; p.fatigue_penalty_factor = 0.5; p.fatigue_threshold_rounds = 3;
    p.stop_candidate_threshold_unacceptable_distance = 1.0; p.stop_candidate_threat_min_vote_share = 0.15; p.stop_candidate_boost_factor = 1.5;
    p.supermajority_threshold = 0.56; p.max_rounds = 20; p.runoff_threshold_rounds = 5; p.runoff_candidate_count = 2;

    int32_t* winner = calloc(n, sizeof(int32_t));
    int32_t* rounds = calloc(n, sizeof(int32_t));
    int64_t* whist = calloc((size_t)E + 1, sizeof(int64_t));
    int64_t* rhist = calloc(21, sizeof(int64_t));
    int64_t totals[2] = {0, 0};

    csim_run_args a;
    memset(&a, 0, sizeof a);
    a.params = &p; a.n_param_sets = 1; a.wiring = wiring; a.engine = engine; a.precision = precision;
    a.sim_id_begin = 1000; a.n_sims = n; a.seed = seed;
    a.winner = winner; a.rounds = rounds; a.winner_hist = whist; a.rounds_hist = rhist; a.totals = totals;
    a.buffers_on_device = 0; a.stream = NULL;
    if (grp) {
        float ms[64];
        CHECK(csim_group_run(grp, &a, ms));
        for (int i = 0; i < n_devices; ++i)
            if (!(ms[i] >= 0.0f)) { fprintf(stderr, "device %d kernel time %f\n", i, ms[i]); return 3; }
    } else {
        CHECK(csim_run(ctx, &a));
    }

    for (uint64_t i = 0; i < n; ++i) printf("%d %d\n", (int)winner[i], (int)rounds[i]);
    printf("totals %lld %lld\n", (long long)totals[0], (long long)totals[1]);
    printf("whist");
    for (int c = 0; c <= E; ++c) printf(" %lld", (long long)whist[c]);
    printf("\nrhist");
    for (int r = 0; r <= 20; ++r) printf(" %lld", (long long)rhist[r]);
    printf("\n");

    /* an unsupported request must come back as a status + message, never as a silent fallback */
    a.precision = CSIM_FP64; a.engine = CSIM_ENGINE_PREFIX;
    csim_status st = grp ? csim_group_run(grp, &a, NULL) : csim_run(ctx, &a);
    printf("prefix+fp64 -> %d (%s)\n", (int)st, st == CSIM_OK ? "" : csim_last_error());

    if (grp) CHECK(csim_group_destroy(grp));
    else CHECK(csim_destroy(ctx));
    free(x); free(region); free(pap); free(winner); free(rounds); free(whist); free(rhist);
    return 0;
}
