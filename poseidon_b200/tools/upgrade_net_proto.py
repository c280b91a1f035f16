"""upgrade_net_proto_text / _binary: rewrite a deprecated (V0 / legacy transform) net definition in the
current schema.  reference: src/caffe/util/upgrade_proto.cpp:598-623 and Caffe's tools/upgrade_net_proto_*.cpp."""
from __future__ import annotations

import sys

from .. import proto as P


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print("usage: upgrade_net_proto IN OUT   (text in -> text out, binary in -> binary out)")
        return 1
    with open(argv[0], "rb") as f:
        raw = f.read()
    try:
        raw.decode("utf-8")
        is_text = True
    except UnicodeDecodeError:
        is_text = False
    net = P.read_net(argv[0])
    if is_text:
        P.write_text(argv[1], net)
    else:
        P.write_binary(argv[1], net)
    return 0


if __name__ == "__main__":
    sys.exit(main())
