#!/usr/bin/env python
"""Command entry point mirroring the reference's manage.py for the one command on the
hot path:  python manage.py run_mcts <n_tiles> <rows> <cols> [--from_file] [--device cuda:0]"""
import sys


def main():
    if len(sys.argv) < 2 or sys.argv[1] != "run_mcts":
        sys.exit("usage: manage.py run_mcts n_tiles rows cols [--n_sim N] [--from_file] "
                 "[--avg_depth] [--profile] [--device cuda:N]")
    from visapi_b200.run_mcts import main as run
    run(sys.argv[2:])


if __name__ == "__main__":
    main()
