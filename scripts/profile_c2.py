import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cProfile, pstats, torch, time
import gnnb200
from gnnb200 import data as data_mod, models, synthetic
dev = torch.device('cuda')
graphs = synthetic.tu_like_graphs('ENZYMES', 128, seed=42)
batch = data_mod.Batch.from_data_list([data_mod.Data(**g) for g in graphs]).to(dev)
ft = models.FinetuneGNN(dev, 'ENZYMES', 'full_finetune'); ft.train()
opt = torch.optim.AdamW(ft.param_groups)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(ft(batch), batch.y)
    loss.backward(); opt.step()
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize(); print('ms/step', (time.perf_counter() - t0) / 50 * 1e3)
# split: forward only / backward / optimizer
def timed(fn, n=50):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
with torch.no_grad():
    print('fwd no_grad ms', timed(lambda: ft(batch)))
print('fwd grad ms', timed(lambda: ft(batch)))
pr = cProfile.Profile(); pr.enable()
for _ in range(30): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('cumulative').print_stats(35)
