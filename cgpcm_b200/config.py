"""Global configuration, mirroring the reference's ``src/config.py:3-4``.

``reg`` is the jitter added by ``reg()`` to every matrix that is factorised
(``src/core/tf_util.py:310-320``); tasks overwrite it at import time (``src/tasks/toy.py:7``: 1e-6,
``ou.py:9``: 1e-5, ``crude.py:9``: 1e-4).  It is read at every evaluation and passed explicitly through
the C-ABI.  ``dtype`` is always float64.
"""
reg = 1e-8
dtype = 'float64'
