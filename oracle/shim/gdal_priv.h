// GDAL stand-in: declarations only, so that the reference's imageop.h parses.  The TIFF writers are
// out of scope; every entry point aborts if it is ever reached.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdlib>
enum GDALRWFlag { GF_Read, GF_Write };
enum GDALDataType { GDT_UInt16 = 2 };
enum CPLErr { CE_None = 0, CE_Failure = 3 };
enum GDALColorInterp { GCI_RedBand, GCI_GreenBand, GCI_BlueBand, GCI_AlphaBand };
struct GDALRasterBand {
    CPLErr RasterIO(GDALRWFlag, int, int, int, int, void *, int, int, GDALDataType, long long, long long) { abort(); }
    void SetColorInterpretation(GDALColorInterp) { abort(); }
};
struct GDALDataset { GDALRasterBand *GetRasterBand(int) { abort(); } };
struct GDALDriver { GDALDataset *Create(const char *, int, int, int, GDALDataType, char **) { abort(); } };
struct GDALDriverManager { GDALDriver *GetDriverByName(const char *) { abort(); } };
inline GDALDriverManager *GetGDALDriverManager() { abort(); }
inline void GDALClose(GDALDataset *) { abort(); }
inline void GDALAllRegister() {}
inline char **CSLParseCommandLine(const char *) { abort(); }
inline char **CSLSetNameValue(char **, const char *, const char *) { abort(); }
inline void CSLDestroy(char **) { abort(); }
