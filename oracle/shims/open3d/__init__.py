"""Import shim (test infrastructure): Open3D is absent in the build container.  The reference touches it in one place on
the hot path's caller side — instance_utils.same_instance (ovo/utils/instance_utils.py:16-22):
`pcd1.compute_point_cloud_distance(pcd2)`, by Open3D's documentation "for each point in the source point cloud, the
distance to the [nearest point of the] target point cloud" — restated here on SciPy's KD-tree in float64 so that the
reference's OVO.update_map can run for the loop-closure golden (oracle/gen_golden.py:gen_update_map)."""
import numpy as np


class _Utility:
    @staticmethod
    def Vector3dVector(a):
        return np.asarray(a, np.float64)


class _PointCloud:
    def __init__(self):
        self.points = np.zeros((0, 3))

    def compute_point_cloud_distance(self, other):
        from scipy.spatial import KDTree
        return KDTree(np.asarray(other.points)).query(np.asarray(self.points), k=1)[0]


class _Geometry:
    PointCloud = _PointCloud


utility = _Utility()
geometry = _Geometry()
