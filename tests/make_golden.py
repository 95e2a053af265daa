#!/usr/bin/env python
"""Extract the reference's golden vectors into tests/golden/*.npz.

The reference's own test fixtures (``/root/reference/test/fixtures.py``:
``dataset_ERA5pressurelevel`` lines 28-1162, ``dataset_soundings`` lines
1165-3441) are xarray ``Dataset.from_dict`` literals.  xarray/dask are not
installed here and the reference tree does not travel to the GPU box, so this
script reads the dict literals with ``ast`` (they contain ``np.nan``), applies
the same ``transpose()`` the soundings fixture applies (fixtures.py:3439), and
stores plain numpy arrays.  It is run ONCE in the build container; its output is
committed.  Nothing at test/bench time reads ``/root/reference``.

    python tests/make_golden.py [/root/reference]
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _dict_literals(path):
    tree = ast.parse(open(path).read())
    out = {}
    for fn in tree.body:
        if not isinstance(fn, ast.FunctionDef):
            continue
        lits = []
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict):
                lits.append(eval(compile(ast.Expression(node.value), path, 'eval'), {'np': np}))
        if lits:
            out[fn.name] = lits
    return out


def main(ref='/root/reference'):
    lits = _dict_literals(os.path.join(ref, 'test', 'fixtures.py'))

    # --- dataset_soundings: stored (level, n2), used transposed (n2, level) ---
    (snd,) = lits['dataset_soundings']
    out = {}
    for name, var in snd['data_vars'].items():
        a = np.asarray(var['data'], dtype=np.float64)
        if var['dims'] == ('level', 'n2'):
            a = np.ascontiguousarray(a.T)
        out[name] = a
    np.savez_compressed(os.path.join(HERE, 'golden', 'ref_soundings.npz'), **out)

    # --- dataset_ERA5pressurelevel: first literal = surface, second = 3d ---
    surf, d3 = lits['dataset_ERA5pressurelevel']
    out = {}
    for name, var in surf['data_vars'].items():
        out['surf_' + name] = np.asarray(var['data'])
    for name, var in d3['data_vars'].items():
        assert var['dims'] == ('longitude', 'level')
        out['lev_' + name] = np.asarray(var['data'], dtype=np.float64)
    out['level'] = np.asarray(d3['coords']['level']['data'], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'golden', 'ref_era5pl.npz'), **out)
    for f in ('ref_soundings.npz', 'ref_era5pl.npz'):
        z = np.load(os.path.join(HERE, 'golden', f))
        print(f, {k: z[k].shape for k in z.files})


if __name__ == '__main__':
    main(*sys.argv[1:])
