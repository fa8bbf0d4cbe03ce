#!/bin/bash
# Device-code fingerprint of a built library:  tools/sass_fingerprint.sh [lib.so]
# md5 of `cuobjdump -sass` with the instruction encodings and the path-dependent anonymous-namespace hashes removed, so
# that two builds of the same device code in different directories compare equal.  Used to check that host-only or
# comment-only commits made after the last GPU validation ship exactly the validated kernels.
lib=${1:-$(dirname "$0")/../mycelium_fea_project_b200/libmycelium_fea_b200.so}
cuobjdump -sass "$lib" | sed 's#/\*[0-9a-f]\{4,\}\*/##g' | sed -E 's/_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_0-9]+_cu_[0-9a-f]+/ANON/g' | md5sum | cut -d' ' -f1
