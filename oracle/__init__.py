"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the D2D-PPO hot path (envs, nets, GAE/returns, PPO
updates) used as the parity checker for the CUDA product in
``d2d-ppo_b200/``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``d2d-ppo_b200/`` imports, links or executes anything here;
the product path fails loudly when its CUDA library is missing.

Parity pinning: the reference ships NO tests, golden vectors or KATs for
this path.  The oracle is therefore pinned against
outputs of the UNMODIFIED reference executed in the build container
(``oracle/gen_golden.py`` imports ``/root/reference`` through
``oracle/ref_harness.py`` and writes ``tests/golden/*.npz``); the committed
fixtures travel to the GPU box, the reference does not.
"""
