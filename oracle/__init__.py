"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference GNAN path ("Oracle").
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
