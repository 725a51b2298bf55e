"""Stand-in for the GDAL python bindings, which are not installed in this image.  TEST INFRASTRUCTURE ONLY.

The reference's Module.py / Loss.py do `from CommonFunc import *`, and CommonFunc.py:17-19 imports
`osgeo.gdal/ogr/osr` at module scope, so the names must exist for the unmodified reference files to import.

`gdal` additionally implements the small slice of the API that the reference's data path calls (data_utils.py:33-39,
104-105, 191-213; Demo_USSS.py:441-448) on top of an in-memory numpy store, so the UNMODIFIED `GDALDataset`,
`NORMALIZE`, `Dataset_meanstd` and `GDALwriteDefault` can run here and generate golden vectors for the raster staging /
tiled writer rows:

    gdal.register("x.tif", array[nband, ysize, xsize])   ->  gdal.Open("x.tif")
    ds.RasterXSize / RasterYSize / RasterCount, ds.GetRasterBand(b).ReadAsArray(x, y, w, h) / .WriteArray(arr, x, y),
    ds.GetDriver().Create(path, xsize, ysize, nband, dtype), Get/SetGeoTransform, Get/SetProjection
"""
import types as _types

import numpy as _np

gdal = _types.ModuleType("osgeo.gdal")
ogr = _types.ModuleType("osgeo.ogr")
osr = _types.ModuleType("osgeo.osr")
gdal.GDT_Float32 = 6
gdal.GDT_Int32 = 5
gdal.GDT_Byte = 1

_STORE = {}


class _Band:
    def __init__(self, arr2d):
        self._a = arr2d

    def ReadAsArray(self, xoff=0, yoff=0, xsize=None, ysize=None):
        xsize = self._a.shape[1] - xoff if xsize is None else xsize
        ysize = self._a.shape[0] - yoff if ysize is None else ysize
        return self._a[yoff:yoff + ysize, xoff:xoff + xsize].copy()

    def WriteArray(self, arr, xoff=0, yoff=0):
        arr = _np.asarray(arr)
        self._a[yoff:yoff + arr.shape[0], xoff:xoff + arr.shape[1]] = arr

    def SetNoDataValue(self, v):
        pass

    def FlushCache(self):
        pass


class _Driver:
    def Create(self, path, xsize, ysize, nband, dtype=6):
        np_dtype = {6: _np.float32, 5: _np.int32, 1: _np.uint8}.get(dtype, _np.float32)
        ds = _Dataset(_np.zeros((nband, ysize, xsize), dtype=np_dtype))
        _STORE[path] = ds
        return ds


class _Dataset:
    def __init__(self, arr):
        self.array = arr
        self.RasterCount, self.RasterYSize, self.RasterXSize = arr.shape
        self._gt, self._proj = (0.0, 1.0, 0.0, 0.0, 0.0, -1.0), ""

    def GetRasterBand(self, b):
        return _Band(self.array[b - 1])

    def GetDriver(self):
        return _Driver()

    def GetGeoTransform(self):
        return self._gt

    def SetGeoTransform(self, gt):
        self._gt = gt

    def GetProjection(self):
        return self._proj

    def SetProjection(self, p):
        self._proj = p

    def FlushCache(self):
        pass


def _register(path, array):
    """Put a [nband, ysize, xsize] array behind `path` (2-D arrays become one band)."""
    array = _np.asarray(array)
    if array.ndim == 2:
        array = array[None]
    _STORE[path] = _Dataset(array)
    return _STORE[path]


def _open(path, *args):
    return _STORE.get(path)


gdal.register = _register
gdal.Open = _open
gdal.store = _STORE
