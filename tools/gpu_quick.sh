#!/bin/bash
# quick GPU check: run the given pytest selection under a hard timeout
timeout ${T:-240} python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -25
