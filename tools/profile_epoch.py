"""Host-side profile of one training epoch (phases 1-3 + eval) on the synthetic baby-shape dataset."""
import cProfile, pstats, io, os, sys, tempfile, time
sys.path.insert(0, '.')
import torch
from diffmm_b200 import Main, synth
from diffmm_b200.Conf import Config
name = sys.argv[1] if len(sys.argv) > 1 else 'baby'
U, I, dims = synth.SHAPES[name]
root = tempfile.mkdtemp(prefix="diffmm_prof_")
synth.write_dataset(root, name, synth.interactions(U, I, seed=0), synth.features(I, dims, seed=0))
os.chdir(root)
cfg = Config(); cfg.data.name = name; cfg.base.precision = 'bf16'; cfg.train.epoch = 3; cfg.train.test_batch = 1024
Main.seed_it(0)
h = Main.DataHandler(cfg); h.LoadData()
coach = Main.Coach(h, cfg); coach.prepareModel()
coach.phase_seconds = {}
coach.trainEpoch(); coach.testEpoch(); torch.cuda.synchronize()
coach.phase_seconds = {}
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
coach.trainEpoch(); res = coach.testEpoch(); torch.cuda.synchronize()
pr.disable()
print("epoch s", time.perf_counter() - t0, coach.phase_seconds, res)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25); print(s.getvalue()[:5000])
