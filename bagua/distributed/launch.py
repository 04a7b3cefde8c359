from bagua_b200.distributed.launch import *  # noqa: F401,F403
from bagua_b200.distributed.launch import main

if __name__ == "__main__":
    import sys

    sys.exit(main())
