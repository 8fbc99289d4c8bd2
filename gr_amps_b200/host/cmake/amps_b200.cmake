# Include from gr-amps's lib/CMakeLists.txt (reference lib/CMakeLists.txt:28-30 builds gnuradio-amps from lib/*.cc):
#   set(AMPS_B200_DIR /path/to/this/repo)
#   include(${AMPS_B200_DIR}/gr_amps_b200/host/cmake/amps_b200.cmake)
# It swaps the block bodies for the ones in gr_amps_b200/host/lib (same class names, make() signatures, ports) and links
# the prebuilt CUDA library; gr-amps itself needs no CUDA language support.
find_library(AMPS_B200_LIB amps_b200 HINTS ${AMPS_B200_DIR}/gr_amps_b200 NO_DEFAULT_PATH)
if(NOT AMPS_B200_LIB)
    message(FATAL_ERROR "libamps_b200.so not found: run `make -C ${AMPS_B200_DIR}/gr_amps_b200` (nvcc, sm_100a) first")
endif()
set(AMPS_B200_SOURCES
    ${AMPS_B200_DIR}/gr_amps_b200/host/lib/blocks_impl.cc
    ${AMPS_B200_DIR}/gr_amps_b200/csrc/proto.cc)
set(AMPS_B200_INCLUDE_DIRS
    ${AMPS_B200_DIR}/include
    ${AMPS_B200_DIR}/gr_amps_b200/host/include)     # no gr_shim: the real <gnuradio/sync_block.h> and <pmt/pmt.h> are used
# usage:
#   add_library(gnuradio-amps SHARED ${AMPS_B200_SOURCES})
#   target_include_directories(gnuradio-amps PRIVATE ${AMPS_B200_INCLUDE_DIRS})
#   target_link_libraries(gnuradio-amps ${Boost_LIBRARIES} ${GNURADIO_ALL_LIBRARIES} ${AMPS_B200_LIB})
install(FILES ${AMPS_B200_DIR}/gr_amps_b200/host/grc/amps_recc_iq.xml ${AMPS_B200_DIR}/gr_amps_b200/host/grc/amps_forward_iq.xml
        DESTINATION share/gnuradio/grc/blocks)
