"""Host-side runtime around the hot path: training-step semantics, synthetic data, one-process-per-GPU
launchers (DDP training with NCCL gradient all-reduce; collective-free batch-sharded inference)."""
