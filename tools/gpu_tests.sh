#!/bin/bash
# the whole GPU suite, verbose about the slow full-size tests
python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25
