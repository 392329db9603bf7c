"""Importable alias of ``distributed-multi-agent-slam-swarm-robotics-system_b200/``.

The package directory name required by the repository layout contains hyphens, which Python
cannot import directly; this shim points its ``__path__`` at that directory so that
``occgrid_b200.dual_bot_mapper`` etc. resolve to the files that live there.  No code lives here.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'distributed-multi-agent-slam-swarm-robotics-system_b200')
__path__ = [_REAL]
with open(_os.path.join(_REAL, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, '__init__.py'), 'exec'))
