"""CLI shim with the reference's entry-point name:  python main.py --dataset netflix [flags of utility/parser.py]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    from llmrec_b200.main import main
    main()
