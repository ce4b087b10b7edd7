"""TEST INFRASTRUCTURE: stand-in for the third-party `pysobol` package (requirements.txt:6 of the reference, not installed
here).  `pysobol.data._sobol_data` is the Joe-Kuo new-joe-kuo-6.21201 table as a flat integer stream; it is re-serialised
from the identical table that ships with SciPy (oracle/sobol_table.py)."""
