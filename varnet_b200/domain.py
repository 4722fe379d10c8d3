"""Spatial domains and their structured discretisation (host side, NumPy only).

Mirror of the reference's `Domain1D`, `PolygonDomain2D` and `Mesh`
(`/root/reference/Domain.py:119-159,216-519,586-696`): same constructors, same
`getMesh(discNum, bDiscNum, rfrac, sortflg, discTol)` outputs.  matplotlib is not a
dependency here: point-in-polygon is a NumPy crossing-number test
(`Domain.py:372-373` used `matplotlib.path.Path.contains_points`); interior grid points
of the supported configurations are strictly inside, where both tests agree.
Plot helpers (`domPlot`, `meshPlot`) are out of scope.
"""
import math

import numpy as np


class Mesh:
    """Record of a discretised domain (Domain.py:119-159)."""

    def __init__(self, dim, dof, coordinates, he, bIndNum, bdof, bCoordinates, discNum=[], bDiscNum=[]):
        self.dim = dim
        self.dof = dof
        self.coordinates = coordinates
        self.he = he
        self.bIndNum = bIndNum
        self.bdof = bdof
        self.bCoordinates = bCoordinates
        self.discNum = discNum
        self.bDiscNum = bDiscNum


class Domain:
    def __init__(self, dim, lim):
        self.dim = dim
        self.lim = np.array(lim)

    def scaleCoord(self, x):
        """Centre and scale coordinates to [-1,1] (Domain.py:71-88)."""
        if np.shape(x)[1] != self.dim:
            raise ValueError('Input dimensions are incompatible with domain dimension!')
        cen = np.mean(self.lim, axis=0)
        scale = np.diff(self.lim, axis=0)
        return (x - cen) / scale * 2

    def isInside(self, x):
        raise Exception('This function must be redefined in the subclass!')

    def getMesh(self):
        raise Exception('This function must be redefined in the subclass!')


def _split_counts(n, rfrac):
    nrand = math.floor(n * rfrac)
    return nrand, n - nrand


def _axis_points(lo, hi, n, rfrac, sortflg, tol):
    """`n` test-function centres in [lo+tol, hi-tol]: a random part then a uniform part."""
    nrand, nuni = _split_counts(n, rfrac)
    parts = [np.linspace(lo + tol, hi - tol, nuni)]
    if nrand > 0:
        parts.insert(0, np.random.uniform(lo + tol, hi - tol, nrand).reshape((nrand,) + parts[0].shape[1:]))
    pts = np.concatenate(parts, axis=0)
    if rfrac > 0 and sortflg:
        pts = np.sort(pts)
    return pts


def points_in_polygon(vertices, pts):
    """Crossing-number test; vertices [n,2] (open ring), pts [m,2] -> bool[m]."""
    v = np.asarray(vertices, dtype=float)
    pts = np.asarray(pts, dtype=float)
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), dtype=bool)
    xj, yj = v[-1]
    for xi, yi in v:
        straddle = (yi > y) != (yj > y)
        with np.errstate(divide='ignore', invalid='ignore'):
            xcross = xi + (y - yi) * (xj - xi) / (yj - yi)
        inside ^= straddle & (x < xcross)
        xj, yj = xi, yi
    return inside


def _shoelace(vertices):
    v = np.asarray(vertices, dtype=float)
    # same pairing as UtilityFunc.polyArea (x<-col 1, y<-col 0), |.|/2
    x, y = v[:, 1], v[:, 0]
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


class Domain1D(Domain):
    """Interval domain (Domain.py:586-696)."""

    def __init__(self, interval=np.array([-1.0, 1.0])):
        interval = np.asarray(interval)
        if interval.ndim != 1:
            raise ValueError('interval must be a vector!')
        super().__init__(1, np.reshape(interval, [2, 1]))
        self.bIndNum = 2
        self.measure = interval[1] - interval[0]

    def isInside(self, x, tol=0.):
        if np.shape(x)[1] != self.dim:
            raise ValueError('Vertex dimensions are incompatible with domain dimension!')
        return (self.lim[0] + tol <= x) * (x <= self.lim[1] - tol)

    def getMesh(self, discNum=100, bDiscNum=None, rfrac=0, sortflg=True, discTol=None):
        if np.size(discNum) != 1:
            raise ValueError('number of discretization points must be a scalar!')
        if np.shape(discNum) != ():
            discNum = discNum[0]
        rfrac = min(max(rfrac, 0), 1)
        lim = self.lim
        he = (lim[1] - lim[0]) / (discNum + 1)               # element size, array of shape (1,)
        tol = he if discTol is None else np.asarray(discTol).item()
        coordinates = _axis_points(lim[0], lim[1], discNum, rfrac, sortflg, tol).reshape(discNum, 1)
        return Mesh(dim=1, dof=discNum, coordinates=coordinates, he=he, bIndNum=2,
                    bdof=np.ones(2, dtype=int), bCoordinates=np.reshape(lim, [2, 1, 1]), discNum=discNum)

    def meshTot(self, discNum=100):
        """Mesh including the end points (Domain.py:699-741)."""
        if np.shape(discNum) != ():
            discNum = discNum[0]
        lim = self.lim
        he = (lim[1] - lim[0]) / (discNum - 1)
        coordinates = np.linspace(lim[0], lim[1], discNum).reshape(discNum, 1)
        return Mesh(dim=1, dof=discNum, coordinates=coordinates, he=he, bIndNum=None, bdof=None,
                    bCoordinates=None, discNum=discNum)


class PolygonDomain2D(Domain):
    """Polygon with optional polygonal obstacles (Domain.py:216-519)."""

    def __init__(self, vertices=np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]]), obsVertices=[]):
        vertices = np.asarray(vertices, dtype=float)
        if vertices.shape[1] != 2:
            raise ValueError('Vertex dimensions are incompatible with domain dimension!')
        if type(obsVertices) is not list:
            raise ValueError('obstacle polygons must be given as a list of matrices!')
        super().__init__(2, np.vstack([vertices.min(axis=0), vertices.max(axis=0)]))
        self.vertexNum = len(vertices)
        self.vertices = vertices
        self.obsNum = len(obsVertices)
        self.obsVertices = obsVertices
        self.bIndNum = len(vertices) + sum(len(o) for o in obsVertices)
        self.boundryGeom = np.concatenate([self.boundaryLims(p) for p in [vertices] + list(obsVertices)], axis=0)
        self.measure = _shoelace(vertices)

    @staticmethod
    def boundaryLims(vertices):
        """[bIndNum,2,2]: the two end points of every edge (Domain.py:272-284)."""
        v = np.asarray(vertices, dtype=float)
        return np.stack([v, np.roll(v, -1, axis=0)], axis=1)

    def isInside(self, x, tol=0.):
        if np.shape(x)[1] != self.dim:
            raise ValueError('Vertex dimensions are incompatible with domain dimension!')
        inside = points_in_polygon(self.vertices, x)
        for obs in self.obsVertices:
            inside &= ~points_in_polygon(obs, x)
        return inside

    def innerDisc(self, discNum, rfrac=0., sortflg=True, discTol=None):
        if np.size(discNum) not in (1, 2):
            raise ValueError('\'discNum\' dimension incompatible!')
        if np.size(discNum) == 1:
            discNum = [discNum, discNum]
        if discTol is not None and np.size(discTol) not in (1, 2):
            raise ValueError('\'discTol\' dimension incompatible!')
        if discTol is not None and np.size(discTol) == 1:
            discTol = [discTol, discTol]
        lim = self.lim
        rf = rfrac ** 0.5                                    # random fraction per dimension
        he, axes = [], []
        for d in range(2):
            n = discNum[d]
            h = (lim[1, d] - lim[0, d]) / (n + 1)
            he.append(h)
            tol = h if discTol is None else np.asarray(discTol[d]).item()
            axes.append(_axis_points(lim[0, d], lim[1, d], n, rf, sortflg, tol))
        ne = int(np.prod(discNum))
        xy = np.hstack([np.tile(axes[0], discNum[1]).reshape(ne, 1),       # x fastest (Domain.py:475-476)
                        np.repeat(axes[1], discNum[0]).reshape(ne, 1)])
        return np.array(he), xy[self.isInside(xy), :]

    @staticmethod
    def boundaryDisc(vertices, bDiscNum, rfrac=0, sortflg=True):
        v = np.asarray(vertices, dtype=float)
        closed = np.vstack([v, v[0, :]])
        edge = np.diff(closed, axis=0)
        length = np.linalg.norm(edge, axis=1)
        bdof, coord = [], []
        for i in range(len(v)):
            n = math.ceil(bDiscNum * length[i])
            bdof.append(n)
            nrand, nuni = _split_counts(n, rfrac)
            step = np.hstack([np.random.uniform(size=nrand), np.linspace(0.0, 1.0, num=nuni)])
            if rfrac > 0 and sortflg:
                step = np.sort(step)
            step = np.tile(step, [2, 1]).T
            coord.append(closed[i, :] + edge[i, :] * step)
        return bdof, coord

    def getMesh(self, discNum=100, bDiscNum=50, rfrac=0, sortflg=True, discTol=None):
        he, coordinates = self.innerDisc(discNum, rfrac, sortflg, discTol)
        bdof, bCoordinates = self.boundaryDisc(self.vertices, bDiscNum, rfrac, sortflg)
        for obs in self.obsVertices:
            d, c = self.boundaryDisc(obs, bDiscNum, rfrac, sortflg)
            bdof.extend(d); bCoordinates.extend(c)
        return Mesh(dim=2, dof=len(coordinates), coordinates=coordinates, he=he, bIndNum=self.bIndNum,
                    bdof=bdof, bCoordinates=bCoordinates, discNum=discNum, bDiscNum=bDiscNum)
