"""CPU ORACLE — test infrastructure, not product code.

Restatement of the reference's hot path (codec_port / entropy_port / rem_port in torch-CPU fp32, rans_port.c in plain C)
plus the recipe that compiles the reference's own native coder into oracle/_ref (build_ref) and the scripts that run the
UNMODIFIED reference to write the fixtures under tests/golden (gen_golden, gen_golden_rem, gen_checkpoint_golden).

Parity status: PINNED.  `python -m oracle.gen_golden --check [--cases headline config3 config4 custmap table800]` shows
the restatement bit-identical to the real reference (streams, reconstructions, likelihoods: max |d| = 0) on four flag
sets x six qualities at the small fixture shapes and at the shapes BASELINE.json names (768x512, 16x3x256x256 with 13
levels, 2048x1408); tests/test_oracle_*.py hold it there, tests/test_oracle_coder.py pins the C coder to the compiled
reference and to the known-answer vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this package; the product
package (progressivecodec_b200) never does, and fails loudly without its CUDA library.
"""
