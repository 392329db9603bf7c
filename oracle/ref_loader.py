"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference modules from ``/root/reference`` (authoring container
only; the GPU box has no ``/root/reference``).  ``dual_bot_mapper`` and
``playback_dual_session`` hard-exit when pygame is missing
(server_nodes/dual_bot_mapper.py:29-34, simulation_tools/playback_dual_session.py:27-32),
so an empty ``pygame`` / ``pygame.gfxdraw`` stub is injected first.  Nothing on the
integration path touches pygame.
"""
import contextlib
import importlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('OCCGRID_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'server_nodes', 'dual_bot_mapper.py'))


def _stub_pygame():
    if 'pygame' not in sys.modules:
        pg = types.ModuleType('pygame')
        pg.gfxdraw = types.ModuleType('pygame.gfxdraw')
        sys.modules['pygame'] = pg
        sys.modules['pygame.gfxdraw'] = pg.gfxdraw


def load_dual_bot_mapper():
    """Returns the reference ``dual_bot_mapper`` module (OccupancyGrid, PoseGraphSLAM, …)."""
    _stub_pygame()
    p = os.path.join(REFERENCE_ROOT, 'server_nodes')
    if p not in sys.path:
        sys.path.insert(0, p)
    return importlib.import_module('dual_bot_mapper')


def load_playback():
    _stub_pygame()
    p = os.path.join(REFERENCE_ROOT, 'simulation_tools')
    if p not in sys.path:
        sys.path.insert(0, p)
    return importlib.import_module('playback_dual_session')


class QuietSLAM:
    """Wraps the reference PoseGraphSLAM, swallowing its per-closure print (:316-318)."""

    def __init__(self, ref_module):
        self._s = ref_module.PoseGraphSLAM()

    def add_pose(self, *a):
        with contextlib.redirect_stdout(io.StringIO()):
            return self._s.add_pose(*a)

    @property
    def closures(self):
        return self._s.closures


# ---------------------------------------------------------------------------------------------
# map_merger.py: rclpy, nav_msgs and open3d are absent here.  The stubs below let the UNMODIFIED
# server_nodes/map_merger.py be imported and EXECUTED: its own NumPy lines (grid_to_pcd :64-85,
# publish_global_map :87-127) and its own callback state machine (map_callback :35-62) run as
# written; only the three Open3D point-cloud methods and the registration call, which live in the
# absent third-party library, delegate to the restatements in oracle/merge_oracle.py and
# oracle/icp_oracle.py (those stay "restated from Open3D's published algorithm").
# ---------------------------------------------------------------------------------------------
class _Obj:
    """Attribute bag: msg.info.origin.position.x = ... works without declaring anything."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        v = _Obj()
        object.__setattr__(self, name, v)
        return v


class _StubLogger:
    def __init__(self):
        self.lines = []

    def info(self, s):
        self.lines.append(('info', s))

    def warn(self, s):
        self.lines.append(('warn', s))


class _StubNode:
    """rclpy.node.Node as far as map_merger.py:9-29,113-127 uses it."""

    def __init__(self, name):
        self._name = name
        self._params = {}
        self._logger = _StubLogger()
        self.published = []

    def declare_parameter(self, name, default):
        self._params.setdefault(name, list(default) if isinstance(default, (list, tuple)) else default)

    def get_parameter(self, name):
        v = self._params[name]
        o = _Obj()
        o.get_parameter_value = lambda: type('PV', (), {'integer_array_value': v})()
        return o

    def get_logger(self):
        return self._logger

    def create_subscription(self, typ, topic, cb, depth):
        return (topic, cb)

    def create_publisher(self, typ, topic, depth):
        node = self
        return type('Pub', (), {'publish': lambda self_, msg: node.published.append(msg)})()

    def get_clock(self):
        now = type('T', (), {'to_msg': lambda self_: 0})()
        return type('C', (), {'now': lambda self_: now})()


class StubPointCloud:
    """open3d.geometry.PointCloud as far as map_merger.py uses it: .points, is_empty(),
    transform(), +=, voxel_down_sample().  The arithmetic of the last three is
    oracle/merge_oracle.py's restatement of Open3D's published algorithm."""

    def __init__(self):
        import numpy as np
        self.points = np.zeros((0, 3))

    def is_empty(self):
        return len(self.points) == 0

    def transform(self, T):
        import numpy as np
        from oracle import merge_oracle as MO
        p = np.asarray(self.points)
        x, y = MO.transform_points(p[:, 0], p[:, 1], T)
        self.points = np.stack([x, y, np.zeros_like(x)], 1)
        return self

    def __iadd__(self, other):
        import numpy as np
        self.points = np.concatenate([np.asarray(self.points), np.asarray(other.points)])
        return self

    def voxel_down_sample(self, voxel_size):
        import numpy as np
        from oracle import merge_oracle as MO
        p = np.asarray(self.points)
        x, y = MO.voxel_down_sample(p[:, 0], p[:, 1], voxel_size)
        out = StubPointCloud()
        out.points = np.stack([x, y, np.zeros_like(x)], 1)
        return out


class RegistrationScript:
    """What the stubbed `registration_icp` returns: either the next (transformation, fitness)
    of a supplied list (configs supply T directly), or — mode 'icp' — the
    restated algorithm of oracle/icp_oracle.py on the clouds the reference hands over."""

    def __init__(self):
        self.queue = []
        self.mode = 'script'
        self.calls = []

    def __call__(self, source, target, threshold, init, estimation, criteria):
        import numpy as np
        assert np.array_equal(np.asarray(init), np.identity(4))
        r = _Obj()
        if self.mode == 'icp':
            from oracle import icp_oracle as IO
            s, t = np.asarray(source.points), np.asarray(target.points)
            T, fit, rmse, its = IO.registration_icp(s[:, 0], s[:, 1], t[:, 0], t[:, 1], threshold,
                                                     criteria.max_iteration)
            r.transformation, r.fitness, r.inlier_rmse = T, fit, rmse
        else:
            T, fit = self.queue.pop(0)
            r.transformation, r.fitness = np.asarray(T, float).reshape(4, 4), fit
        self.calls.append((len(source.points), len(target.points), float(threshold), r.fitness))
        return r


def load_map_merger():
    """Returns (module, registration_script): the UNMODIFIED reference `map_merger` module
    imported under rclpy / nav_msgs / open3d stubs."""
    import numpy as np
    script = RegistrationScript()
    rclpy = types.ModuleType('rclpy')
    rclpy.init = lambda args=None: None
    rclpy.spin = lambda node: None
    rclpy.shutdown = lambda: None
    rclpy_node = types.ModuleType('rclpy.node')
    rclpy_node.Node = _StubNode
    rclpy.node = rclpy_node
    nav = types.ModuleType('nav_msgs')
    nav_msg = types.ModuleType('nav_msgs.msg')
    nav_msg.OccupancyGrid = type('OccupancyGrid', (_Obj,), {})
    nav_msg.MapMetaData = type('MapMetaData', (_Obj,), {})
    nav.msg = nav_msg
    o3d = types.ModuleType('open3d')
    o3d.geometry = types.ModuleType('open3d.geometry')
    o3d.geometry.PointCloud = StubPointCloud
    o3d.utility = types.ModuleType('open3d.utility')
    o3d.utility.Vector3dVector = lambda pts: np.array(pts, dtype=np.float64).reshape(-1, 3)
    o3d.pipelines = types.ModuleType('open3d.pipelines')
    reg = types.ModuleType('open3d.pipelines.registration')
    reg.registration_icp = script
    reg.TransformationEstimationPointToPoint = lambda: 'point_to_point'
    reg.ICPConvergenceCriteria = lambda max_iteration=30: type('Crit', (), {'max_iteration': max_iteration})()
    o3d.pipelines.registration = reg
    for name, mod in (('rclpy', rclpy), ('rclpy.node', rclpy_node), ('nav_msgs', nav), ('nav_msgs.msg', nav_msg),
                      ('open3d', o3d), ('open3d.geometry', o3d.geometry), ('open3d.utility', o3d.utility),
                      ('open3d.pipelines', o3d.pipelines), ('open3d.pipelines.registration', reg)):
        sys.modules.setdefault(name, mod)
    p = os.path.join(REFERENCE_ROOT, 'server_nodes')
    if p not in sys.path:
        sys.path.insert(0, p)
    mod = importlib.import_module('map_merger')
    # the script the module's `o3d` actually points at (a second load reuses the first stubs)
    return mod, mod.o3d.pipelines.registration.registration_icp


def make_ref_grid_msg(mod, data, width, height, res, ox, oy, frame_id='map'):
    """nav_msgs/OccupancyGrid as map_merger.py:64-71 reads it (msg.data is a Python list, as
    rclpy delivers it)."""
    msg = mod.OccupancyGrid()
    msg.header.frame_id = frame_id
    msg.info.width, msg.info.height, msg.info.resolution = int(width), int(height), float(res)
    msg.info.origin.position.x, msg.info.origin.position.y = float(ox), float(oy)
    msg.data = [int(v) for v in data]
    return msg
