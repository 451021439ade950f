#!/usr/bin/env python
"""Drop-in entry point: `python G2Vec.py EXPRESSION_FILE CLINICAL_FILE NETWORK_FILE RESULT_NAME [options]`
(same command line as mathcom/G2Vec; the work is in g2vec_b200/cli.py)."""
from g2vec_b200.cli import main

if __name__ == "__main__":
    main()
