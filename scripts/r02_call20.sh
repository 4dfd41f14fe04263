#!/bin/bash
# persistent-lanes experiment (SWK_ASYNC=1; kernel removed again after this measurement: profiles/r02_async_lanes_ab.log, profiles/README.md)
echo "the kernel variant this script measured was removed; see profiles/README.md 'Tried and rejected this round'"
