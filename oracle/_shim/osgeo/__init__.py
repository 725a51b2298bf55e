"""Stand-in for the GDAL python bindings, which are not installed in this image.

The reference's Module.py / Loss.py do `from CommonFunc import *`, and CommonFunc.py:17-19 imports
`osgeo.gdal/ogr/osr` at module scope.  The hot path never touches GDAL, so empty namespaces are enough to
import the unmodified reference files.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
"""
import types as _types

gdal = _types.ModuleType("osgeo.gdal")
ogr = _types.ModuleType("osgeo.ogr")
osr = _types.ModuleType("osgeo.osr")
gdal.GDT_Float32 = 6
gdal.GDT_Int32 = 5
