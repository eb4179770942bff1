"""Builds ``deciphon_core._cffi`` -- the cffi API-mode extension python-core imports its ``ffi`` and
``lib`` from (python-core/deciphon_core/{scan,batch,press,error}.py: ``from deciphon_core._cffi
import ffi, lib``) -- against libdeciphon_b200.so instead of the reference's static libdeciphon.

python-core/build_ext.py:88-112 does the same with ``#include "deciphon.h"`` and the libraries
deciphon, h3client, h3result, hmmer_reader, imm, lio, lite_pack; here the one library is
deciphon_b200 and the header is include/deciphon_b200.h, which declares the same ``dcp_*`` entry
points.  The cdef below is the interface python-core binds (python-core/deciphon_core/interface.h:
press, scan, batch, dcp_error_string, ``extern "Python" void callback(void *)``) restated from
include/deciphon_b200.h, plus this library's dcpb200_* extensions.

    python -m deciphon_b200.compat.build_cffi [target-dir]

writes ``<target-dir>/deciphon_core/_cffi.<abi>.so`` (default target: deciphon_b200/compat/_site).
Put ``<target-dir>`` ahead of python-core's own build on sys.path, or copy the module over the one
python-core built, and python-core's Scan / Batch / PressContext classes run unmodified on the GPU.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)

CDEF = """
struct dcp_press;
struct dcp_scan;
struct dcp_batch;

struct dcp_press *dcp_press_new(void);
int               dcp_press_setup(struct dcp_press *, int gencode_id, float epsilon);
int               dcp_press_open(struct dcp_press *, char const *hmm, char const *db);
long              dcp_press_nproteins(struct dcp_press const *);
int               dcp_press_next(struct dcp_press *);
bool              dcp_press_end(struct dcp_press const *);
int               dcp_press_close(struct dcp_press *);
void              dcp_press_del(struct dcp_press const *);

struct dcp_scan *dcp_scan_new(void);
void             dcp_scan_del(struct dcp_scan const *);
int              dcp_scan_setup(struct dcp_scan *, char const *dbfile, int port, int num_threads,
                                bool multi_hits, bool hmmer3_compat, bool cache,
                                void (*callback)(void *), void *userdata);
int              dcp_scan_run(struct dcp_scan *, struct dcp_batch *, char const *product_dir);
void             dcp_scan_interrupt(struct dcp_scan *);
int              dcp_scan_progress(struct dcp_scan const *);

struct dcp_batch *dcp_batch_new(void);
void              dcp_batch_del(struct dcp_batch *);
int               dcp_batch_add(struct dcp_batch *, long id, char const *name, char const *data);
void              dcp_batch_reset(struct dcp_batch *);

char const *dcp_error_string(int error_code);

int    dcpb200_db_info(char const *dbfile, int *num_proteins, long *total_core_size, float *epsilon);
int    dcpb200_scan_num_gpus(struct dcp_scan const *);
int    dcpb200_scan_num_shards(struct dcp_scan const *);
double dcpb200_scan_counter(struct dcp_scan const *, int what);

FILE *fopen(char const *filename, char const *mode);
FILE *fdopen(int, char const *);
int   fclose(FILE *);

extern "Python" void callback(void *);
"""


def build(target: str | None = None, verbose: bool = False) -> str:
    from cffi import FFI
    target = os.path.abspath(target or os.path.join(HERE, "_site"))
    os.makedirs(os.path.join(target, "deciphon_core"), exist_ok=True)
    ffibuilder = FFI()
    ffibuilder.cdef(CDEF)
    ffibuilder.set_source(
        "deciphon_core._cffi",
        '#include <stdio.h>\n#include "deciphon_b200.h"\n',
        language="c",
        libraries=["deciphon_b200"],
        library_dirs=[PKG],
        include_dirs=[os.path.join(ROOT, "include")],
        extra_link_args=["-Wl,-rpath," + PKG],
    )
    return ffibuilder.compile(tmpdir=target, verbose=verbose)


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else None, verbose=True))
