"""PPO / IPPO learners and baseline policies (reference: algorithms/*.py) on sm_100a kernels."""
