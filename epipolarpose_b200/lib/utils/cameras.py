"""Camera projection-matrix conventions -- mirror of the part of the reference
lib/utils/cameras.py on the hot path (:5-17,120-131,149-150): P = K.[R | R.(-T)]
in float64, X_cam = R.(X - T).  Distortion / h5 loading is dataset preparation
(out of scope)."""
import numpy as np


class Camera():
    def __init__(self, cam_params):
        self.cam_params = cam_params
        self.R, self.T, self.f, self.c, self.k, self.p, self.name = cam_params
        self.camera_matrix = self.get_intrinsic_matrix()
        self.tvec = self.get_tvec()
        self.projection_matrix = self.get_projection_matrix()

    def get_intrinsic_matrix(self):
        fx, fy = np.asarray(self.f).reshape(-1)[:2]
        cx, cy = np.asarray(self.c).reshape(-1)[:2]
        return np.array([[fx, 0., cx], [0., fy, cy], [0., 0., 1.]]).astype(np.double)

    def get_tvec(self):
        return np.dot(self.R, np.negative(self.T))

    def get_disp_matrix(self):
        return np.concatenate((self.R, self.get_tvec()), axis=1)

    def get_projection_matrix(self):
        T = self.tvec
        if len(T.shape) < 2:
            T = np.expand_dims(T, axis=-1)
        return np.dot(self.get_intrinsic_matrix(), np.concatenate((self.R, T), axis=1))
