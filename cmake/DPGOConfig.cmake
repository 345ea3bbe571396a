# DPGOConfig.cmake -- lets dpgo_ros's own build find the B200 path under the package name it already asks for:
#   find_package(DPGO REQUIRED)            (CMakeLists.txt:6 of mit-acl/dpgo_ros)
#   target_link_libraries(... DPGO ...)    (CMakeLists.txt:151-154)
# Point CMake at this directory (-DDPGO_DIR=<repo>/cmake) after building the library
# (python -c "import __graft_entry__ as g; g.build()").  The imported target carries the shim headers
# (include/DPGO/*.h, namespace DPGO) and the C ABI header (include/dpgo_b200.h).
get_filename_component(_DPGO_B200_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(DPGO_INCLUDE_DIRS "${_DPGO_B200_ROOT}/include")
set(DPGO_LIBRARY "${_DPGO_B200_ROOT}/dpgo_ros_b200/libdpgo_b200.so")
if(NOT EXISTS "${DPGO_LIBRARY}")
  message(FATAL_ERROR "DPGO (B200 path): ${DPGO_LIBRARY} is missing -- build it first (there is no CPU fallback)")
endif()
if(NOT TARGET DPGO)
  add_library(DPGO SHARED IMPORTED)
  set_target_properties(DPGO PROPERTIES
    IMPORTED_LOCATION "${DPGO_LIBRARY}"
    INTERFACE_INCLUDE_DIRECTORIES "${DPGO_INCLUDE_DIRS}"
    INTERFACE_COMPILE_FEATURES cxx_std_17)
endif()
set(DPGO_LIBRARIES DPGO)
set(DPGO_FOUND TRUE)
