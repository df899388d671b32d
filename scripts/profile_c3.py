import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cProfile, pstats, torch, time, random, importlib.util
spec = importlib.util.spec_from_file_location('bench', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'bench.py'))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
import gnnb200
from gnnb200 import data as data_mod, models, tasks as task_mod
dev = torch.device('cuda')
lists = b._graph_lists()
torch.manual_seed(0)
pm = models.PretrainableGNN(dev, b.S4_DOMAINS, b.S4_TASKS); pm.train()
popt = torch.optim.AdamW(pm.parameters(), lr=1e-4)
temp, grl = task_mod.TemperatureScheduler(1000), task_mod.GRLScheduler(50, 100)
tasks = task_mod.instantiate_tasks(pm, b.S4_TASKS, grl, temp)
batches = {d: b._make_batch(data_mod, lists[d], dev) for d in b.S4_DOMAINS}
gen = torch.Generator().manual_seed(42); random.seed(42)
def run_task(name):
    popt.zero_grad(set_to_none=True)
    loss, _ = tasks[name].compute_loss(batches, gen); loss.backward()
for name in tasks:
    for _ in range(2): run_task(name)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): run_task(name)
    torch.cuda.synchronize(); print(f'{name:16s} {(time.perf_counter()-t)/5*1e3:8.1f} ms (loss+backward)')
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    for name in tasks: run_task(name)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
