"""TEST INFRASTRUCTURE ONLY.  CPU restatements of the reference's algorithms, used as checkers by tests/, by
bench.py's cpu_baseline / --impl reference legs (and the cpu_baseline leg of tools/bench_search_fits.py) and by
__graft_entry__.smoke(); nothing in the product package imports this directory.

* npp_oracle.py    -- encoding -> MLP forward -> sigmoid + masked MSE -> backward -> Adam (pinned by tests/golden/golden_{encoding,topk,top1,relu}.npz);
                      search stage: encode_search, forward_light / backward_light / train_step_light
                      (pinned by tests/golden/golden_light.npz)
* robust_oracle.py -- the adaptive robust pixel loss and its gradients (pinned by tests/golden/golden_robust.npz)
"""
