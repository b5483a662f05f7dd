"""The three tables BoneLengthLoss reads from the reference's mesh/bone_length.py (:36-56) — data only.

The reference module computes nothing at import that these tables depend on, but importing it loads the mesh, builds a
HandSynthesizer on the GPU and imports matplotlib/cv2/dataset; the mirror carries the tables alone.
The same 35 pairs and rest lengths are compiled into csrc/pose_losses.cu."""
joint_1 = [3, 2, 3, 8, 2, 2, 9] + [8, 4, 8, 7, 4, 6] + [7, 0, 5, 7, 7, 6, 6]
joint_2 = [2, 9, 8, 2, 4, 10, 10] + [4, 10, 7, 4, 6, 10] + [6, 5, 1, 0, 5, 5, 1]
for _idx in range(5):       # three bones per finger (mesh/bone_length.py:47-54)
    joint_1 += [11 + _idx * 6, 13 + _idx * 6, 15 + _idx * 6]
    joint_2 += [12 + _idx * 6, 14 + _idx * 6, 16 + _idx * 6]

uniform_length = [25.212656021118164, 18.249488830566406, 27.5742244720459, 38.532264709472656, 25.10819435119629,
                  31.173757553100586, 18.329626083374023, 19.15080451965332, 16.209327697753906, 21.52261734008789,
                  32.740535736083984, 30.58920669555664, 33.205970764160156, 11.672294616699219, 17.084707260131836,
                  17.084720611572266, 16.697546005249023, 23.92103385925293, 20.87999725341797, 22.58038330078125,
                  27.55999755859375, 15.471183776855469, 13.214692115783691, 21.748210906982422, 13.021653175354004,
                  16.643720626831055, 18.83765983581543, 12.724685668945312, 16.238431930541992, 18.04928970336914,
                  11.045844078063965, 11.320968627929688, 30.078536987304688, 16.255985260009766, 19.434825897216797]
