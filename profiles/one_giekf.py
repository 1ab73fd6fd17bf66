import sys
sys.path.insert(0, "/root/repo/profiles")
import measure_giekf as m
m.run(int(sys.argv[1]), 2)
