from oracle.sobol_table import joe_kuo_stream

_sobol_data = joe_kuo_stream(21201)
