"""xcape_b200 — B200-native column kernels behind xcape's ``calc_cape`` / ``calc_srh`` API.

Mirrors the reference package layout (``xcape/__init__.py:5`` exports ``core``): the public
functions are ``xcape_b200.core.calc_cape`` and ``xcape_b200.core.calc_srh`` with the
reference's signatures plus ``method='cuda'``.
"""
__version__ = '0.1.0'
__all__ = ['core']
