"""Vehicle/obstacle shapes: only what the NLP needs -- checkpoints + radii,
canvas limits and room hyperplanes (reference omgtools/basics/shape.py:62-67,
146-171, 330-336, 350-362).  Drawing is out of scope."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/basics/shape.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np


class Shape(object):
    def __init__(self, dimensions):
        self.n_dim = dimensions


class Shape2D(Shape):
    def __init__(self):
        Shape.__init__(self, 2)

    @staticmethod
    def rotate(orientation, coordinate):
        if isinstance(orientation, np.ndarray):
            orientation = orientation.reshape(-1)[0]
        cth, sth = np.cos(orientation), np.sin(orientation)
        return np.array([[cth, -sth], [sth, cth]]).dot(coordinate)


class Circle(Shape2D):
    def __init__(self, radius):
        Shape2D.__init__(self)
        self.radius = radius
        self.n_chck = 1

    def get_checkpoints(self):
        return [[0., 0.]], [self.radius]

    def get_canvas_limits(self):
        return [np.array([-self.radius, self.radius]),
                np.array([-self.radius, self.radius])]


class Polyhedron(Shape2D):
    def __init__(self, vertices, orientation=0., radius=1e-3):
        Shape2D.__init__(self)
        self.n_vert = vertices.shape[1]
        self.orientation = orientation
        self.vertices = self.rotate(orientation, vertices)
        # radius > 0 is required for polyhedron-polyhedron avoidance
        self.radius = radius

    def get_checkpoints(self):
        chck = [[self.vertices[0, l], self.vertices[1, l]]
                for l in range(self.n_vert)]
        return chck, [self.radius for _ in range(self.n_vert)]

    def get_canvas_limits(self):
        hi, lo = np.amax(self.vertices, axis=1), np.amin(self.vertices, axis=1)
        return [np.array([lo[0], hi[0]]), np.array([lo[1], hi[1]])]

    def get_hyperplanes(self, **kwargs):
        pos = kwargs.get('position', [0, 0])
        v = np.hstack((self.vertices, np.vstack(self.vertices[:, 0])))
        hyperplanes = {}
        for k in range(v.shape[1] - 1):
            d = [v[0][k + 1] - v[0][k], v[1][k + 1] - v[1][k]]
            normal = np.array([-d[1], d[0]]) / np.sqrt(d[0]**2 + d[1]**2)
            b = normal[0] * (v[0][k + 1] + pos[0]) + normal[1] * (v[1][k + 1] + pos[1])
            hyperplanes[k] = {'a': normal, 'b': b}
        return hyperplanes


class Beam(Polyhedron):
    def __init__(self, width, height, orientation=0.):
        self.width, self.height = width, height
        Polyhedron.__init__(self, np.c_[[0.5 * width, 0.], [-0.5 * width, 0.]],
                            orientation=orientation, radius=0.5 * height)


def _tangent_polygon(normals_angle, offsets):
    """Vertices of the polygon with faces a_l . x = b_l, a_l = (sin, cos)(l*dth)."""
    n = len(offsets)
    A = np.array([[np.sin(th), np.cos(th)] for th in normals_angle])
    vertices = np.zeros((2, n))
    for l in range(n):
        a = np.vstack((A[l], A[(l + 1) % n]))
        b = np.array([offsets[l], offsets[(l + 1) % n]])
        vertices[:, l] = np.linalg.solve(a, b)
    return vertices


class RegularPolyhedron(Polyhedron):
    def __init__(self, radius, n_vert, orientation=0.):
        dth = 2 * np.pi / n_vert
        off = [radius * np.cos(np.pi / n_vert)] * n_vert
        vertices = _tangent_polygon([l * dth for l in range(n_vert)], off)
        Polyhedron.__init__(self, vertices, orientation)
        self.radius_outer = radius


class Rectangle(Polyhedron):
    def __init__(self, width, height, orientation=0.):
        self.width, self.height = width, height
        off = [0.5 * height, 0.5 * width, 0.5 * height, 0.5 * width]
        vertices = _tangent_polygon([l * 0.5 * np.pi for l in range(4)], off)
        Polyhedron.__init__(self, vertices, orientation)


class Square(Rectangle):
    def __init__(self, side, orientation=0.):
        Rectangle.__init__(self, side, side, orientation)


class Shape3D(Shape):
    def __init__(self):
        Shape.__init__(self, 3)

    @staticmethod
    def rotate(orientation, coordinate):
        """roll-pitch-yaw rotation (reference shape.py:269-279)."""
        if len(orientation) != 3:
            raise ValueError('Orientation is a list with 3 elements: roll, pitch, yaw!')
        roll, pitch, yaw = orientation
        cpsi, spsi = np.cos(yaw), np.sin(yaw)
        cth, sth = np.cos(pitch), np.sin(pitch)
        cphi, sphi = np.cos(roll), np.sin(roll)
        rot = np.array([[cth*cpsi, sphi*sth*cpsi-cphi*spsi, cphi*sth*cpsi+sphi*spsi],
                        [cth*spsi, sphi*sth*spsi+cphi*cpsi, cphi*sth*spsi-sphi*cpsi],
                        [-sth, sphi*cth, cphi*cth]])
        return rot.dot(coordinate)


class Sphere(Shape3D):
    def __init__(self, radius):
        Shape3D.__init__(self)
        self.radius = radius
        self.n_chck = 1

    def get_checkpoints(self):
        return [[0., 0., 0.]], [self.radius]

    def get_canvas_limits(self):
        return [np.array([-self.radius, self.radius]) for _ in range(3)]


class Polyhedron3D(Shape3D):
    def __init__(self, vertices, orientation=(0, 0, 0), radius=1e-3):
        Shape3D.__init__(self)
        self.n_vert = vertices.shape[1]
        self.radius = radius
        self.orientation = list(orientation)
        self.vertices = self.rotate(self.orientation, vertices)

    def get_checkpoints(self):
        chck = [[self.vertices[0, l], self.vertices[1, l], self.vertices[2, l]]
                for l in range(self.n_vert)]
        return chck, [self.radius for _ in range(self.n_vert)]

    def get_canvas_limits(self):
        hi, lo = np.amax(self.vertices, axis=1), np.amin(self.vertices, axis=1)
        return [np.array([lo[k], hi[k]]) for k in range(3)]


class Cuboid(Polyhedron3D):
    """Vertex order of the reference (shape.py:402-428): 4 bottom, 4 top."""

    def __init__(self, width, depth, height, orientation=(0, 0, 0)):
        self.width, self.depth, self.height = width, depth, height
        base = _tangent_polygon([l * 0.5 * np.pi for l in range(4)],
                                [0.5 * depth, 0.5 * width, 0.5 * depth, 0.5 * width])
        vertices = np.zeros((3, 8))
        vertices[:2, :4] = base
        vertices[2, :4] = -0.5 * height
        vertices[:2, 4:] = base
        vertices[2, 4:] = 0.5 * height
        Polyhedron3D.__init__(self, vertices, orientation)


class RegularPrisma(Polyhedron3D):
    """Prism over a regular polygon; ``radius`` is the circle through the
    vertices (reference shape.py:364-390): n bottom vertices, then n top."""

    def __init__(self, radius, height, n_faces, orientation=(0, 0, 0)):
        self.radius_outer, self.height, self.n_faces = radius, height, n_faces
        dth = 2 * np.pi / n_faces
        base = _tangent_polygon([l * dth for l in range(n_faces)],
                                [radius * np.cos(np.pi / n_faces)] * n_faces)
        vertices = np.zeros((3, 2 * n_faces))
        vertices[:2, :n_faces] = base
        vertices[2, :n_faces] = -0.5 * height
        vertices[:2, n_faces:] = base
        vertices[2, n_faces:] = 0.5 * height
        Polyhedron3D.__init__(self, vertices, orientation)


class Cube(Cuboid):
    def __init__(self, side, orientation=(0, 0, 0)):
        Cuboid.__init__(self, side, side, side, orientation)


class Plate(Polyhedron3D):
    """Flat polygon with half-thickness as radius (reference shape.py:447-454)."""

    def __init__(self, shape2d, height, orientation=(0, 0, 0)):
        self.shape2d = shape2d
        vertices = np.r_[shape2d.vertices, np.zeros((1, shape2d.vertices.shape[1]))]
        Polyhedron3D.__init__(self, vertices, orientation, 0.5 * height)
