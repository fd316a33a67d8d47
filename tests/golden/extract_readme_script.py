#!/usr/bin/env python
"""tests/golden/extract_readme_script.py -- writes tests/golden/readme_script.py: the Python code block of the reference's README
(/root/reference/README.md:40-105, "Then run the following code"), byte for byte.  The fixture lets the GPU box (which has no /root/reference) run
the README script with its imports untouched against the top-level `l2f` / `foundation_policy` packages (tests/test_readme_loop.py)."""
import os
import re

README = "/root/reference/README.md"
HERE = os.path.dirname(os.path.abspath(__file__))

text = open(README).read()
m = re.search(r"Then run the following code:\n```python\n(.*?)\n```", text, re.S)
assert m, "README code block not found"
code = m.group(1) + "\n"
assert "from l2f import vector8 as vector" in code and "from foundation_policy import Raptor" in code
open(os.path.join(HERE, "readme_script.py"), "w").write(code)
print("wrote readme_script.py: %d lines" % code.count("\n"))
