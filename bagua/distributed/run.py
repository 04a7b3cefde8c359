from bagua_b200.distributed.run import *  # noqa: F401,F403
from bagua_b200.distributed.run import main

if __name__ == "__main__":
    main()
