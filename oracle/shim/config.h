// Stand-in for the reference's CMake-generated config.h (cmake/config.h.in:5-9).
#define EXBLAS_VERSION_MAJOR 1
#define EXBLAS_VERSION_MINOR 0
#define EXBLAS_SOURCE_DIR "/root/reference"
#define EXBLAS_BINARY_DIR "/root/reference"
#define USE_EXBLAS
