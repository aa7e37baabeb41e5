"""TEST INFRASTRUCTURE ONLY (oracle). Render the reference's Config.template into a Config.h.

Does what reference src/build.py:25-35 does (string.Template over deepgo/config.py's constants),
without running the reference's build system. Output goes to oracle/_ref/gen/Config.h.
usage: gen_config.py <reference_root> <out_path>
"""
import runpy
import string
import sys
from pathlib import Path

ref, out = Path(sys.argv[1]), Path(sys.argv[2])
consts = runpy.run_path(str(ref / 'src' / 'deepgo' / 'config.py'))
text = (ref / 'src' / 'deepgo' / 'native' / 'cpp' / 'Config.template').read_text()
out.parent.mkdir(parents=True, exist_ok=True)
out.write_text(string.Template(text).substitute(consts))
