"""Drop-in `aphantasia` package: the reference's module names, backed by aphantasia_b200 (put dropin/ on PYTHONPATH)."""
