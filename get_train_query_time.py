"""Drop-in CLI with the reference's script name: `python get_train_query_time.py <dataset> <timestep>`
(/root/reference/get_train_query_time.py:45-58) -> resources/<dataset>_train_query_time.pt, bit-identical.
The implementation lives in rag4dyg_b200/query_time.py."""
from rag4dyg_b200.query_time import *  # noqa: F401,F403
from rag4dyg_b200.query_time import main

if __name__ == "__main__":
    main()
