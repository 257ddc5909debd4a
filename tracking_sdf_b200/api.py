"""Host-side mirror of the reference's two classes, same method names and argument meaning.

    reference (C++, /root/reference/src)                      here
    SDF(m, width, height, depth, origin, delta, eps)          SDF(...)                      sdf.h:78-79
    CameraTracking(max_iter, max_twist_diff, v_h, w_h, sdf)   CameraTracking(...)           camera_tracking.cpp:3-4
    camera_tracking->camera_info_cb(msg)                      .camera_info_cb(K)            camera_tracking.cpp:22-36
    camera_tracking->set_camera_transformation(rot, trans)    .set_camera_transformation()  camera_tracking.cpp:59-65
    camera_tracking->estimate_new_position(sdf, cloud)        .estimate_new_position(sdf, depth)   camera_tracking.cpp:66-245
    sdf->update(camera_tracking, cloud, normals)              .update(camera_tracking, depth)      sdf.cpp:224-315
    sdf->interpolate_distance(voxel_pt, ok)                   .interpolate_distance(pts)    sdf.cpp:127-163
    sdf->update(...) colour part (the cloud's r,g,b)          .update(camera_tracking, depth, rgb) sdf.cpp:294-304
    sdf->interpolate_color(global_coords, color)              .interpolate_color(pts)       sdf.cpp:164-217
    mc->performReconstruction(cloud) + marker fill            .mesh()                       marching_cubes_sdf.cpp:243-287, sdf.cpp:327-385
    rot / trans / rot_inv / rot_inv_trans / K / isKFilled     same attribute names (properties)

The two reference objects share state through raw pointers (the tracker reads the grid, the
volume reads the tracker's pose and K); here both are views of ONE device handle.  Inputs are
depth images instead of PCL clouds: back-projection and normals (done upstream of the reference
by ROS/PCL) are part of the device path.  The C++ twin of this file is
include/tracking_sdf_b200.hpp.
"""
import numpy as np

from . import capi


class SDF:
    def __init__(self, m=256, width=6.0, height=6.0, depth=3.5, sdf_origin=(-3.0, -3.0, -0.5),
                 distance_delta=0.3, distance_epsilon=0.025, **device_opts):
        self._cfg_kw = dict(m=m, width=width, height=height, depth=depth, origin=sdf_origin,
                            distance_delta=distance_delta, distance_epsilon=distance_epsilon, **device_opts)
        self._t = None            # the handle is created when the tracker (which completes the config) attaches
        self.m = m
        self.m_div_width = np.float32(m) / np.float32(width)
        self.m_div_height = np.float32(m) / np.float32(height)
        self.m_div_depth = np.float32(m) / np.float32(depth)

    def _handle(self):
        if self._t is None:
            raise capi.TsdfError(1, "SDF is not attached to a CameraTracking yet (construct CameraTracking(…, sdf))")
        return self._t

    def get_number_of_voxels(self):
        return self.m ** 3

    def get_array_index(self, voxel_coordinates):
        i, j, k = (int(x) for x in voxel_coordinates)
        return self._handle().get_array_index(i, j, k)

    def get_voxel_coordinates(self, x):
        if np.ndim(x) == 0:
            return self._handle().get_voxel_coordinates_idx(int(x))
        return self._handle().get_voxel_coordinates(x)

    def get_global_coordinates(self, voxel_coordinates):
        return self._handle().get_global_coordinates(voxel_coordinates)

    def interpolate_distance(self, voxel_points):
        """-> (values float32 [n], is_interpolated bool [n])"""
        return self._handle().interpolate_distance(voxel_points)

    def update(self, camera_tracking, depth, rgb=None):
        """sdf.cpp:224-315: integrate `depth` (and, when the registered (h, w, 3) uint8 colour image is given,
        the colour running mean of :294-304) at camera_tracking's current pose; -> #voxels updated."""
        if not camera_tracking.isKFilled:
            raise capi.TsdfError(2, "Camera Matrix not received")   # the reference exit(0)s, sdf.cpp:227-229
        if rgb is not None:
            return self._handle().fuse_rgb(depth, rgb)
        return self._handle().fuse(depth)

    def interpolate_color(self, global_points):
        """sdf.cpp:164-217 for n WORLD points -> (n, 4) float32 r,g,b,a."""
        return self._handle().interpolate_color(global_points)

    def mesh(self, iso_level=0.0, colors=False):
        """The visualisation thread's product (sdf.cpp:327-385): marching cubes on the device ->
        (marker points [n, 3] float64 = vertices + sdf_origin[, rgba [n, 4]]); three vertices per triangle."""
        out = self._handle().mesh(iso_level, world=True, colors=colors)
        return out[1:] if colors else out[1]

    @property
    def Color(self):
        """(Color_W, R, G, B) in the reference's layout (sdf.h:49-52)."""
        return self._handle().download_color(capi.LAYOUT_REFERENCE)

    @property
    def D(self):
        return self._handle().download(capi.LAYOUT_REFERENCE)[0]

    @property
    def W(self):
        return self._handle().download(capi.LAYOUT_REFERENCE)[1]


class CameraTracking:
    def __init__(self, gauss_newton_max_iteration=20, maximum_twist_diff=0.001, v_h=1.0, w_h=0.01, sdf=None,
                 image_width=640, image_height=480, pixel_stride=3, metric=capi.POINT_TO_PLANE):
        if sdf is None:
            raise capi.TsdfError(1, "CameraTracking needs the SDF it tracks against")
        cfg = capi.default_config(gauss_newton_max_iteration=gauss_newton_max_iteration,
                                  maximum_twist_diff=maximum_twist_diff, v_h=v_h, w_h=w_h,
                                  image_width=image_width, image_height=image_height,
                                  pixel_stride=pixel_stride, metric=metric, **sdf._cfg_kw)
        self._t = capi.Tsdf(cfg)
        sdf._t = self._t
        self.isKFilled = False
        self._K = np.zeros((3, 3))
        self.last_stats = None

    def camera_info_cb(self, K):
        self._K = np.asarray(K, np.float64).reshape(3, 3).copy()
        self._t.set_intrinsics(self._K)
        self.isKFilled = True

    @property
    def K(self):
        return self._K

    def set_camera_transformation(self, rot, trans):
        self._t.set_pose(rot, trans)

    @property
    def rot(self):
        return self._t.get_pose()[0]

    @property
    def trans(self):
        return self._t.get_pose()[1]

    @property
    def rot_inv(self):
        return self._t.get_pose_inv()[0]

    @property
    def rot_inv_trans(self):
        return self._t.get_pose_inv()[1]

    def project_camera_to_image_plane(self, camera_point):
        ij = self._K @ np.asarray(camera_point, np.float64)
        return np.array([ij[0] / ij[2], ij[1] / ij[2]])

    def project_world_to_camera(self, world_point):
        Ri, ti = self._t.get_pose_inv()
        return Ri @ np.asarray(world_point, np.float64) + ti

    def project_camera_to_world(self, camera_point):
        R, t = self._t.get_pose()
        return R @ np.asarray(camera_point, np.float64) + t

    def estimate_new_position(self, sdf, depth):
        """camera_tracking.cpp:66-245.  Mutates rot/trans like the reference; also returns them."""
        R, t, st = self._t.track(depth)
        self.last_stats = st
        return R, t
