import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import test_learner_gpu as T
dev = torch.device("cuda", 0)
bad = 0
for rep in range(int(sys.argv[1])):
    for cfg in [(32, 4, 136, 9, 11, 4, True), (64, 1, 40, 6, 30, 8, True), (64, 6, 300, 5, 30, 8, True), (48, 3, 64, 7, 20, 8, True), (64, 5, 260, 4, 40, 8, False), (32, 3, 77, 6, 9, 4, False), (64, 6, 256, 5, 30, 8, True), (32, 4, 128, 7, 11, 4, True), (64, 3, 384, 3, 30, 1, True)]:
        try:
            T.test_tensor_core_training_path_vs_autograd(*cfg, dev)
        except AssertionError as e:
            bad += 1; print(rep, cfg, str(e)[:200])
print("bad", bad)
