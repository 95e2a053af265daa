"""Where does the host time of one end-to-end call go outside the C call?  cProfile over 20 calls."""
import cProfile, pstats, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings
d = make_soundings('C2', winds=False)
pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in ('t', 'td', 'ps', 'ts', 'tds')}
for k in pin: pin[k].numpy()[...] = d[k]
hp = {k: v.numpy() for k, v in pin.items()}
f = lambda: cape(d['p'], hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], 1, None, 2, 500., 1, 500., 2)
for _ in range(3): f()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): f()
pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(12)
