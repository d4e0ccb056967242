"""B200-native self-play engine for AlphaSnake-Zero (hot path only: lockstep tic, plane encode, MCTS, value net).

The product is libasz_b200.so (hand-written sm_100a CUDA behind the C ABI of include/asz_b200.h); this package is
the thin Python mirror of the reference's classes (code/utils/*.py) on top of it."""
__version__ = "0.1.0"
