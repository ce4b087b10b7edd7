import sys

from . import PATH, run_script

if len(sys.argv) < 2:
    sys.exit(__doc__ or 'usage: python -m ptina_b200.compat [--path | script.py [args...]]')
if sys.argv[1] == '--path':
    print(PATH)
else:
    run_script(sys.argv[1], sys.argv[2:])
