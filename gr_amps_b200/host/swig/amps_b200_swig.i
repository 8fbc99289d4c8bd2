/* -*- c++ -*- */
/* The two lines a gr-amps maintainer appends to swig/amps_swig.i (reference swig/amps_swig.i:19-28) for the new blocks;
 * the five existing blocks keep their entries because class names and make() signatures are unchanged. */
%{
#include "amps/recc_iq.h"
#include "amps/forward_iq.h"
%}
%include "amps/recc_iq.h"
GR_SWIG_BLOCK_MAGIC2(amps, recc_iq);
%include "amps/forward_iq.h"
GR_SWIG_BLOCK_MAGIC2(amps, forward_iq);
